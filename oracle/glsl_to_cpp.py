#!/usr/bin/env python
"""oracle/glsl_to_cpp.py -- lexical GLSL-compute -> C++ rewrite.  TEST INFRASTRUCTURE ONLY.

    python oracle/glsl_to_cpp.py /root/reference/shaders/PathTracer.comp oracle/_ref/PathTracer.comp.inc

Reads one of the reference's compute shaders WHERE IT LIES and writes a C++ fragment (into the git-ignored
oracle/_ref/) that oracle/ref_shaders_host.cpp #includes inside `namespace glsl { namespace <shader> { ... } }`,
with oracle/glsl_compat.h supplying the GLSL types and built-ins.  Nothing of the shader's logic is rewritten:
expressions, statements, control flow, constants and function bodies pass through token for token.  The rewrite only
touches what is GLSL *syntax* rather than C++:

  #version / local_size layout      -> dropped / `static const uvec3 gl_WorkGroupSize(x, y, z);`
  layout(...) uniform Block {..} n; -> `struct Block_block {..}; static Block_block n;`   (bool -> bool32: 4 bytes in std140)
  layout(...) uniform sampler3D n;  -> `static sampler3D n;`   (likewise image2D / image3D, `writeonly` dropped)
  layout(...) buffer B { T n[]; };  -> `static T* n;`   (storage buffers of PPM_PT.comp: one unsized array each)
  in / out / inout parameters       -> by value / by reference
  lowp / mediump / highp            -> dropped (evaluated as fp32, as crv_oracle.h states)
  1.5  2.0  3.14159265359  1e-3     -> suffixed with `f` (an unsuffixed GLSL literal is a 32-bit float, a C++ one a double)
  v.xyz / v.xy                      -> v.xyz() / v.xy()   (read swizzles; the shaders write no swizzles)
  vec3(f(), g(), h())               -> vec3{f(), g(), h()} when more than one argument draws a random number: GLSL evaluates
                                       constructor arguments left to right, C++ only does so inside braces
  mutable globals (uint rng_state)  -> `static thread_local` (private per invocation in GLSL)
  void main()                       -> void shader_main()

The script prints the uniform blocks and opaque uniforms it found (name, binding) so the host file can be checked
against the reference's descriptor layout (RenderTechniquePT.cpp:20-35, RenderTechniqueSV.cpp:13-22).
"""
import re
import sys

SCALARS = r"(?:uint|int|float|bool|double|vec[234]|ivec[234]|uvec[234]|mat4)"


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def suffix_float_literals(src: str) -> str:
    # decimal floating literals without a suffix: 1.0  1.  .5  3.402823466e+38  1e5
    pat = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
    return pat.sub(lambda m: m.group(1) + "f", src)


def order_constructor_arguments(src: str, log) -> str:
    """vecN( ... ) whose argument list calls generateRandomNumber() more than once -> vecN{ ... } (sequenced left to right)"""
    out, i = [], 0
    for m in re.finditer(r"\b(vec[234])\s*\(", src):
        if m.start() < i:
            continue
        depth, j = 1, m.end()
        while depth and j < len(src):
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        args = src[m.end():j - 1]
        if args.count("generateRandomNumber()") > 1 and not re.search(r"\bvec[234]\s*\(", args):
            log(f"ordered constructor: {m.group(1)}{{{args.strip()[:60]}...}}")
            out.append(src[i:m.start()] + m.group(1) + "{" + args + "}")
            i = j
    out.append(src[i:])
    return "".join(out)


def rewrite(src: str, log) -> str:
    src = strip_comments(src)
    src = re.sub(r"^\s*#version[^\n]*", "", src, flags=re.M)
    src = re.sub(r"^\s*#extension[^\n]*", "", src, flags=re.M)

    # work-group size
    m = re.search(r"layout\s*\(\s*local_size_x\s*=\s*(\d+)\s*,\s*local_size_y\s*=\s*(\d+)\s*,\s*local_size_z\s*=\s*(\d+)\s*\)\s*in\s*;", src)
    if not m:
        raise SystemExit("no local_size layout found")
    src = src[:m.start()] + f"static const uvec3 gl_WorkGroupSize({m.group(1)}u, {m.group(2)}u, {m.group(3)}u);" + src[m.end():]
    log(f"local_size {m.group(1)} {m.group(2)} {m.group(3)}")

    # uniform blocks (UBOs and the push-constant block)
    def block(mm):
        quals, name, body, inst = mm.group(1), mm.group(2), mm.group(3), mm.group(4)
        body = re.sub(r"\bbool\b", "bool32", body)
        log(f"block {name} instance {inst} layout({quals.strip()})")
        return f"struct {name}_block {{{body}}};\nstatic {name}_block {inst};"

    src = re.sub(r"layout\s*\(([^)]*)\)\s*uniform\s+(\w+)\s*\{([^}]*)\}\s*(\w+)\s*;", block, src)

    # opaque uniforms
    def opaque(mm):
        quals, typ, name = mm.group(1), mm.group(2), mm.group(3)
        log(f"opaque {typ} {name} layout({quals.strip()})")
        return f"static {typ} {name};"

    src = re.sub(r"layout\s*\(([^)]*)\)\s*uniform\s+(?:writeonly\s+|readonly\s+)?(sampler3D|image2D|image3D)\s+(\w+)\s*;", opaque, src)
    # storage buffers: one unsized array per block, no instance name
    def ssbo(mm):
        quals, name, typ, arr = mm.group(1), mm.group(2), mm.group(3), mm.group(4)
        log(f"buffer {name}: {typ} {arr}[] layout({quals.strip()})")
        return f"static {typ}* {arr};"

    src = re.sub(r"layout\s*\(([^)]*)\)\s*(?:restrict\s+|readonly\s+|writeonly\s+|coherent\s+)*buffer\s+(\w+)\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", ssbo, src)
    if re.search(r"\blayout\s*\(", src):
        raise SystemExit("unhandled layout qualifier left in the source")

    # parameter qualifiers
    src = re.sub(r"\b(?:out|inout)\s+(const\s+)?(\w+)\s+(\w+)", lambda mm: f"{mm.group(2)}& {mm.group(3)}", src)
    src = re.sub(r"\bin\s+(const\s+)?(\w+)\s+(\w+)", lambda mm: f"{mm.group(1) or ''}{mm.group(2)} {mm.group(3)}", src)
    # precision qualifiers
    src = re.sub(r"\b(?:lowp|mediump|highp)\s+", "", src)
    # literals, swizzles, entry point
    src = suffix_float_literals(src)
    src = re.sub(r"\.(xyz|xy|zxy|yzx)\b(?!\s*\()", r".\1()", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    src = order_constructor_arguments(src, log)

    # mutable globals -> thread_local (brace depth 0 only)
    out, depth = [], 0
    for line in src.split("\n"):
        if depth == 0 and re.match(rf"^\s*{SCALARS}\s+\w+\s*(=[^;{{}}]*)?;\s*$", line):
            log(f"per-invocation global: {line.strip()}")
            line = "static thread_local " + line.lstrip()
        depth += line.count("{") - line.count("}")
        out.append(line)
    return "\n".join(out)


def main():
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    src = open(sys.argv[1]).read()
    notes = []
    cpp = rewrite(src, notes.append)
    with open(sys.argv[2], "w") as f:
        f.write(f"// GENERATED from {sys.argv[1]} by oracle/glsl_to_cpp.py -- do not commit (oracle/_ref is git-ignored)\n")
        f.write(cpp)
        f.write("\n")
    for n in notes:
        print(f"[glsl_to_cpp] {n}")


if __name__ == "__main__":
    main()
