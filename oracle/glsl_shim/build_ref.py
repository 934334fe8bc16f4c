"""Build oracle/_ref/libvgi_refshaders.so: the reference's OWN shader text compiled for the CPU.

TEST INFRASTRUCTURE (part of oracle/). The shaders are read where they lie under
<reference>/VFS/Shaders (default /root/reference), rewritten by the purely syntactic rules R1..R11 below,
wrapped between oracle/glsl_shim/glsl_shim.h (GLSL types, built-ins and fixed-function image / sampler
semantics) and a small driver (oracle/glsl_shim/drivers/<shader>.inc: sets the uniforms the reference's host
code sets and loops over the invocations of its dispatch / draw) and piped into g++ on stdin. No reference
source, rewritten or not, is written anywhere: only the shared library lands in oracle/_ref/ (git-ignored).

Rewrite rules (none of them touches an expression's operators, operands or order):
  R1  comments, `#version`, `#extension` removed; `#include "x.glsl"` inlined (the include guards stay).
  R2  `layout ( ... )` qualifiers removed; the bare `in;` / `out;` left over from `layout(local_size...) in;` etc. removed;
      `layout(constant_id = n) const T X = v;` -> `T X = v;` (a specialisation constant: the driver may override it;
      it stays `const` when it sizes an array); memory qualifiers (readonly, writeonly, coherent, volatile, restrict) removed.
  R3  interface blocks: `uniform|buffer|in|out Name { members } inst;` -> `struct Name { members } inst;`,
      without an instance name -> the members become globals; unsized arrays `T a[];` -> `T* a;`;
      a geometry shader's per-vertex input block `} gs_in[];` -> `gs_in[3]` (triangles).
  R4  global `in T x;` / `out T x;` -> `thread_local T x;`; global `uniform T x;` -> `T x;`.
  R5  parameter qualifiers: `out|inout T x` -> `T& x` (arrays: qualifier dropped, C++ arrays decay), `in T x` -> `T x`.
  R6  float literals get an `f` suffix (GLSL literals are binary32).
  R7  multi-component swizzles `.xyz` -> `.xyz()` (glsl_vec_ops.inl).
  R8  array constructors `T[n]( ... )` -> `{ ... }`.
  R9  `main` -> `shader_main` (macro), `discard` -> flag + return (macro), built-in variables (macros;
      `gl_in[i].gl_Position` -> the member the macro cannot name).
  R10 Q6: radianceDownSample.comp declares `lerpFactor` inside the `if` and uses it after it, so
      the shader does not compile as shipped (tests/test_ref_shaders.py checks that it really fails); the repaired
      build hoists the declaration in front of the `if`, initialised to 0.0 as opacityDownSample.comp:59 does.
  R11 `const` at namespace scope with a braced array initialiser is kept as is (valid C++).
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = os.environ.get("VGI_REFERENCE_ROOT", "/root/reference")
SHADER_DIR = os.path.join(REF_ROOT, "VFS", "Shaders")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
OUT_SO = os.path.join(OUT_DIR, "libvgi_refshaders.so")
# the image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp (same remark as oracle/Makefile): use the system g++
GXX = os.environ.get("VGI_ORACLE_CXX", "/usr/bin/g++")
CXXFLAGS = ["-std=c++20", "-O3", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
            "-Werror=float-conversion", "-Wno-attributes", "-I", HERE]

# shader file -> driver (oracle/glsl_shim/drivers/*.inc)
UNITS = [
    "opacityDownSample.comp", "radianceDownSample.comp", "borderWrapping.comp", "clipmapCleaning.comp",
    "copyAlphaImage.comp", "octreeNodeInit.comp", "octreeNodeFlag.comp", "octreeNodeAlloc.comp",
    "octreeNodeModifyArg.comp", "octreeNodeLeafWrite.comp", "octreeNodeMipmapWrite.comp",
    "voxelConeTracing.frag", "voxelConeTracing_Octree.frag", "specularFilter.frag", "msaaInjectRadiance.frag",
    "msaaVoxelizer.geom", "msaaVoxelizer.frag", "voxelizer.frag",
    "gBufferPass.frag", "msaaVoxelizer.vert", "voxelizer.vert", "voxelizer.geom",
]


def reference_available():
    return os.path.isdir(SHADER_DIR)


def _strip_comments(s):
    s = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def _inline_includes(s, depth=0):
    def repl(m):
        with open(os.path.join(SHADER_DIR, m.group(1)), encoding="utf-8", errors="replace") as f:
            return "\n" + _inline_includes(_strip_comments(f.read()), depth + 1) + "\n"
    assert depth < 8
    return re.sub(r'^[ \t]*#[ \t]*include[ \t]+"([^"]+)"[^\n]*$', repl, s, flags=re.M)


def _block(m):
    kind, name, body, inst = m.group(2), m.group(3), m.group(4), m.group(5)
    body = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", body)     # unsized array -> pointer
    if inst:
        tl = "thread_local " if kind in ("in", "out") else ""
        arr = "[3]" if m.group(6) else ""      # per-vertex input array of a geometry shader fed with triangles
        return f"struct {name} {{{body}}}; {tl}{name} {inst}{arr};"
    return body


def _array_ctor(s):
    # `= T[n]( ... )` -> `= { ... }` with balanced parentheses
    out, pos = [], 0
    for m in re.finditer(r"=\s*\w+\s*\[\s*\d*\s*\]\s*\(", s):
        if m.start() < pos:
            continue
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(s[i], 0)
            i += 1
        out.append(s[pos:m.start()] + "= {" + s[m.end():i - 1] + "}")
        pos = i
    out.append(s[pos:])
    return "".join(out)


def translate(shader, repair_q6=True):
    """The rewritten text of one shader (a Python string; never written to disk)."""
    with open(os.path.join(SHADER_DIR, shader), encoding="utf-8", errors="replace") as f:
        s = _strip_comments(f.read())                                                      # R1
    s = _inline_includes(s)
    s = re.sub(r"^[ \t]*#[ \t]*(version|extension)[^\n]*$", "", s, flags=re.M)
    if shader == "radianceDownSample.comp" and repair_q6:                                   # R10
        s, n = re.subn(r"(\n[ \t]*)(if\s*\(any\(greaterThanEqual\(distanceToCenter)", r"\1float lerpFactor = 0.0;\1\2", s)
        assert n == 1
        s, n = re.subn(r"float(\s+lerpFactor\s*=\s*max\()", r"\1", s)
        assert n == 1
    def _spec(m):      # R2: specialisation constants -> globals the driver may set, unless they size an array
        name = re.match(r"\s*\w+\s+(\w+)", m.group(1)).group(1)
        return ("const" if re.search(r"\[\s*" + name + r"\s*\]", s) else "") + m.group(1)
    s = re.sub(r"\blayout\s*\(\s*constant_id[^)]*\)\s*const\b([^;]*)", _spec, s)
    s = re.sub(r"\b(?:readonly|writeonly|coherent|restrict|volatile)\b", "", s)   # memory qualifiers mean nothing here
    s = re.sub(r"\blayout\s*\([^)]*\)", "", s)
    s = re.sub(r"^\s*(?:in|out)\s*;", "", s, flags=re.M)
    s = re.sub(r"()\b(uniform|buffer|in|out)\s+(\w+)\s*\{([^}]*)\}\s*(\w*)\s*(\[\s*\d*\s*\])?\s*;", _block, s, flags=re.S)    # R3
    s = re.sub(r"^[ \t]*(?:flat\s+)?(?:in|out)\s+(\w+)\s+(\w+)\s*;", r"thread_local \1 \2;", s, flags=re.M)   # R4
    s = re.sub(r"^[ \t]*uniform\s+", "", s, flags=re.M)
    s = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)(\s*\[)", r"\1 \2\3", s)                    # R5
    s = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", s)
    s = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", s)
    s = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])", r"\1f", s)        # R6
    s = re.sub(r"(\bgl_in\s*\[[^\]]*\]\s*\.)gl_Position\b", r"\1cur_position", s)             # R9: gl_Position is a macro
    s = _array_ctor(s)                                                                      # R8
    s = re.sub(r"\.([xyzw]{2,4}|[rgba]{2,4})\b(?!\s*\()", r".\1()", s)                      # R7
    return s


def unit_source(shader, repair_q6=True):
    stem = shader.replace(".", "_")
    with open(os.path.join(HERE, "drivers", stem + ".inc")) as f:
        driver = f.read()
    return ("#include \"glsl_shim.h\"\n#undef M_PI\n#define main shader_main\n"                   # R9
            "namespace {\nGLSL_USING_BUILTINS\n" + translate(shader, repair_q6) + "\n}\n#undef main\n" + driver)


def compile_unit(src, obj):
    return subprocess.run([GXX] + CXXFLAGS + ["-x", "c++", "-", "-c", "-o", obj], input=src, text=True,
                          capture_output=True)


def up_to_date():
    if not os.path.exists(OUT_SO):
        return False
    t = os.path.getmtime(OUT_SO)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE)] + \
           [os.path.join(HERE, "drivers", f) for f in os.listdir(os.path.join(HERE, "drivers"))]
    if reference_available():
        deps += [os.path.join(SHADER_DIR, u) for u in UNITS]
    return all(os.path.getmtime(d) <= t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Returns the path of the library, or None when the reference tree is not present (GPU box: the
    prebuilt library travels with the snapshot)."""
    if not reference_available():
        return OUT_SO if os.path.exists(OUT_SO) else None
    if not force and up_to_date():
        return OUT_SO
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for u in UNITS:
            obj = os.path.join(tmp, u.replace(".", "_") + ".o")
            r = compile_unit(unit_source(u), obj)
            if r.returncode != 0:
                raise RuntimeError(f"{u}: g++ failed\n{r.stderr[-6000:]}")
            if verbose and r.stderr:
                print(r.stderr, file=sys.stderr)
            objs.append(obj)
        subprocess.check_call([GXX, "-shared", "-fopenmp", "-o", OUT_SO] + objs)
    return OUT_SO


if __name__ == "__main__":
    if "--show" in sys.argv:      # development aid: print the rewritten text of one shader
        print(unit_source(sys.argv[sys.argv.index("--show") + 1]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
