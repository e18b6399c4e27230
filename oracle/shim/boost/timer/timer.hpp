// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <chrono>
#include <cstdint>
namespace boost { namespace timer {
typedef std::int_least64_t nanosecond_type;
struct cpu_times { nanosecond_type wall, user, system; void clear() { wall = user = system = 0; } };
class cpu_timer {
	std::chrono::steady_clock::time_point m_t0;
	cpu_times m_acc;
	bool m_stopped;
public:
	cpu_timer() { start(); }
	void start() { m_acc.clear(); m_stopped = false; m_t0 = std::chrono::steady_clock::now(); }
	void stop() { if(!m_stopped) { m_acc = elapsed(); m_stopped = true; } }
	void resume() { if(m_stopped) { cpu_times c = m_acc; start(); m_t0 -= std::chrono::nanoseconds(c.wall); } }
	bool is_stopped() const { return m_stopped; }
	cpu_times elapsed() const
	{
		if(m_stopped) return m_acc;
		cpu_times c;
		c.wall = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - m_t0).count();
		c.user = c.wall; c.system = 0;
		return c;
	}
};
} }
