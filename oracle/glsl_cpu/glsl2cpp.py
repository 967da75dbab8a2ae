#!/usr/bin/env python
"""oracle/glsl_cpu/glsl2cpp.py -- TEST INFRASTRUCTURE.

Lexical adapter GLSL ES 3.00 -> C++ for the reference's fragment shaders, run at BUILD time by oracle/Makefile:

    python oracle/glsl_cpu/glsl2cpp.py /root/reference/shader/tracer.fs oracle/_ref/tracer.gen.inc
    python oracle/glsl_cpu/glsl2cpp.py /root/reference/texture_packer.js oracle/_ref/blit.gen.inc --js-template fsStr

(the second form takes the GLSL out of a JavaScript template string, ``let fsStr = `...`;``: the atlas blit shader of
texture_packer.js:103-121 lives there)

The output (git-ignored, never committed: it is derived from the reference's source) is #included inside a namespace
behind glsl_body.inc and compiled by g++.  Nothing about the shader's logic is touched -- no statement is added,
removed or reordered; the line count is preserved so that compiler diagnostics point at the shader's own lines.
Only surface syntax that C++ spells differently is rewritten:

  1. `#version`, `precision ...;` and comments are blanked;
  2. global `in` / `out` / `uniform` declarations and other mutable globals become `static thread_local` variables
     (one fragment per thread); `uniform T name[N];` becomes `uniform_array<T> name;`
  3. parameter qualifiers: `in T x` -> `T x`, `out T x` / `inout T x` -> `T& x`;
  4. floating literals get an `f` suffix (GLSL literals are binary32; a C++ `1.0` is a double and would change the
     arithmetic);
  5. constructor calls with two or more arguments, `T(a, b, ...)` for vector / matrix types and the shader's own
     structs, become `T{a, b, ...}`: GLSL evaluates arguments left to right (GLSL ES 3.00 section 6.1.1), C++ only
     guarantees that order for braced lists -- tracer.fs:426 calls rnd() in two arguments of one constructor;
  6. `main` is renamed `shader_main`.
"""
import re
import sys

VEC_TYPES = ["vec2", "vec3", "vec4", "ivec2", "uvec2", "uvec4", "mat3"]
GLOBAL_TYPES = r"(?:float|int|uint|bool|vec[234]|ivec[234]|uvec[234]|sampler2D|sampler2DArray)"


def blank_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith("//", i):
            j = src.find("\n", i)
            j = n if j < 0 else j
            i = j
        elif src.startswith("/*", i):
            j = src.find("*/", i)
            j = n if j < 0 else j + 2
            out.append("".join(c if c == "\n" else " " for c in src[i:j]))
            i = j
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def brace_constructors(src, types):
    """T( ... , ... ) -> T{ ... , ... } when T is a constructor type and the call has >= 2 top-level arguments."""
    s = list(src)
    for m in re.finditer(r"\b(%s)\s*\(" % "|".join(map(re.escape, types)), src):
        before = src[:m.start()].rstrip()
        # `vec3 name(` is a function declaration, `T(` after an identifier-free position is a constructor call; a
        # struct's own definition never has `(` right after its name
        open_pos = m.end() - 1
        depth, commas, j = 0, 0, open_pos
        while j < len(src):
            c = src[j]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    break
            elif c == "," and depth == 1:
                commas += 1
            j += 1
        if commas >= 1 and not re.search(r"\b(struct)$", before):
            s[open_pos], s[j] = "{", "}"
    return "".join(s)


def convert(src):
    src = blank_comments(src)
    structs = re.findall(r"\bstruct\s+(\w+)", src)
    lines = []
    for line in src.split("\n"):
        st = line.strip()
        if st.startswith("#version") or st.startswith("precision "):
            lines.append("")
            continue
        m = re.match(r"^uniform\s+(\w+)\s+(\w+)\s*\[\s*\w+\s*\]\s*;(.*)$", line)
        if m:
            lines.append("static thread_local uniform_array<%s> %s;%s" % m.groups())
            continue
        m = re.match(r"^(?:uniform|in|out)\s+(%s\s+\w+(?:\s*\[\s*\w+\s*\])?\s*;.*)$" % GLOBAL_TYPES, line)
        if m:
            lines.append("static thread_local " + m.group(1))
            continue
        if re.match(r"^%s\s+\w+\s*;" % GLOBAL_TYPES, line):  # e.g. `float seed;` at global scope
            lines.append("static thread_local " + line)
            continue
        lines.append(line)
    src = "\n".join(lines)
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", src)
    src = brace_constructors(src, VEC_TYPES + structs)
    src = re.sub(r"\bmain\b", "shader_main", src)
    return src


if __name__ == "__main__":
    text = open(sys.argv[1]).read()
    if "--js-template" in sys.argv:
        name = sys.argv[sys.argv.index("--js-template") + 1]
        m = re.search(r"\blet\s+%s\s*=\s*`(.*?)`\s*;" % re.escape(name), text, re.S)
        assert m, "no template string %s in %s" % (name, sys.argv[1])
        body = m.group(1).split("\n")          # the string's lines share the JavaScript code's indentation: remove it
        ind = min(len(l) - len(l.lstrip(" ")) for l in body[1:] if l.strip())
        body = [body[0]] + [l[ind:] if l.strip() else "" for l in body[1:]]
        text = "\n" * text[:m.start(1)].count("\n") + "\n".join(body)   # keep the JavaScript file's line numbers
    out = convert(text)
    assert out.count("\n") == text.count("\n")
    with open(sys.argv[2], "w") as f:
        f.write("/* GENERATED from %s by oracle/glsl_cpu/glsl2cpp.py -- do not commit */\n" % sys.argv[1])
        f.write("#line 1 \"%s\"\n" % sys.argv[1])
        f.write(out)
