#!/usr/bin/env python3
"""Generate link-time stubs for the in-place build of the reference hider.  TEST INFRASTRUCTURE ONLY.

    gen_ref_stubs.py <reference root> <output dir>      (output dir = oracle/_ref/gen, git-ignored)

oracle/ref_hider.cpp drives the reference's OWN hider sources (libs/core/bucketprocessor.cpp,
imagepixel.cpp, micropolygon.cpp, imagebuffer.cpp, occlusion.cpp, ...) compiled in place.  Those
sources talk to two big interfaces whose implementations (renderer.cpp, the shader VM) pull in the
whole renderer: CqRenderer (libs/core/renderer.h) and IqShaderData
(include/aqsis/shadervm/ishaderdata.h).  This script reads the DECLARATIONS of their virtual
functions from the reference headers and emits do-nothing definitions ("called a stub: abort") so
that the vtables exist; the handful of members the hider really calls are written by hand in
ref_hider.cpp and listed in HAND below.  Nothing is copied into the repository: the output lives
under oracle/_ref/.
"""
import os
import re
import sys

# defined by hand in ref_hider.cpp
HAND = {"~CqRenderer", "poptCurrent", "Time", "GetIntegerOption", "GetFloatOption", "Initialise"}


def strip_defaults(args):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return ", ".join(re.sub(r"\s*=\s*[^,]+$", "", a.strip()) for a in out)


def renderer_stubs(ref):
    src = open(os.path.join(ref, "libs/core/renderer.h")).read()
    body = src[src.index("class CqRenderer : public IqRenderer"):]
    out = []
    for m in re.finditer(r"^\s*virtual\s+([^;{}()]*?)([~\w]+)\s*\(([^;{}]*)\)\s*(const)?\s*;\s*$", body, re.M):
        ret, name, args, const = m.group(1).strip(), m.group(2), m.group(3), m.group(4) or ""
        if name in HAND:
            continue
        out.append(f"{ret} CqRenderer::{name}({strip_defaults(args)}) {const} {{ refStubAbort(\"CqRenderer::{name}\"); }}")
    return "\n".join(out) + "\n"


def shaderdata_stubs(ref):
    src = open(os.path.join(ref, "include/aqsis/shadervm/ishaderdata.h")).read()
    out = []
    for d in re.findall(r"virtual\s+([^;{}]*?)\s*=\s*0\s*;", src, re.S):
        d = " ".join(d.split())
        m = re.match(r"(.*?)(\w+)\s*\((.*)\)\s*(const)?$", d)
        ret, name, args, const = m.group(1).strip(), m.group(2), m.group(3), m.group(4) or ""
        # the pointer getters for points and colours are the only members the hider uses
        if name in ("GetPointPtr", "GetColorPtr"):
            continue
        out.append(f"virtual {ret} {name}({strip_defaults(args)}) {const} {{ refStubAbort(\"IqShaderData::{name}\"); }}")
    return "\n".join(out) + "\n"


def attributes_stubs(ref):
    """IqAttributes (include/aqsis/core/iattributes.h): every pure virtual but GetStringAttribute, which the trim-curve
    hit test reads ("trimcurve" "sense") and ref_hider.cpp answers."""
    src = open(os.path.join(ref, "include/aqsis/core/iattributes.h")).read()
    out = []
    for d in re.findall(r"virtual\s+([^;{}]*?)\s*=\s*0\s*;", src, re.S):
        d = " ".join(d.split())
        m = re.match(r"(.*?)(\w+)\s*\((.*)\)\s*(const)?$", d)
        ret, name, args, const = m.group(1).strip(), m.group(2), m.group(3), m.group(4) or ""
        if name == "GetStringAttribute":
            continue
        out.append(f"virtual {ret} {name}({strip_defaults(args)}) {const} {{ refStubAbort(\"IqAttributes::{name}\"); }}")
    return "\n".join(out) + "\n"


def main():
    ref, outdir = sys.argv[1], sys.argv[2]
    os.makedirs(outdir, exist_ok=True)
    open(os.path.join(outdir, "attributes_stubs.inc"), "w").write(attributes_stubs(ref))
    open(os.path.join(outdir, "renderer_stubs.inc"), "w").write(renderer_stubs(ref))
    open(os.path.join(outdir, "shaderdata_stubs.inc"), "w").write(shaderdata_stubs(ref))


if __name__ == "__main__":
    main()
