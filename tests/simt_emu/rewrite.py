#!/usr/bin/env python3
"""tests/simt_emu/rewrite.py -- TEST INFRASTRUCTURE ONLY.

Turns the three CUDA-only constructs g++ cannot parse into calls of the SIMT interpreter; everything else in the kernel
sources is compiled unchanged against include/cuda_runtime.h:

  kernel<T...><<<grid, block, smem, stream>>>(args);   ->  emu::launch(grid, block, smem, stream, [&] { kernel<T...>(args); });
  extern __shared__ T name[];                          ->  T *name = reinterpret_cast<T *>(emu::dynamic_smem());
  asm volatile("bar.sync ...") / ("prefetch...")       ->  emu::named_barrier(id, count) / nothing

usage: rewrite.py <input .cu/.cuh/.h> <output>
"""
import re
import sys


def match_back_angle(text, end):
    """text[end-1] == '>': index of the matching '<'."""
    depth = 0
    i = end - 1
    while i >= 0:
        c = text[i]
        if c == '>':
            depth += 1
        elif c == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def match_forward(text, start, open_c, close_c):
    depth = 0
    i = start
    while i < len(text):
        c = text[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + open_c)


def split_top_level(text):
    parts, depth, cur = [], 0, ""
    for c in text:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def rewrite_launches(text):
    out, pos = "", 0
    while True:
        at = text.find("<<<", pos)
        if at < 0:
            return out + text[pos:]
        # kernel expression: identifier with optional template arguments, directly before <<<
        name_end = at
        j = at
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == '>':
            j = match_back_angle(text, j)
        k = j
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        kernel = text[k:name_end].strip()
        close = text.index(">>>", at)
        config = split_top_level(text[at + 3:close])
        while len(config) < 4:
            config.append("0" if len(config) == 2 else "nullptr")
        paren = text.index("(", close)
        paren_end = match_forward(text, paren, "(", ")")
        semi = text.index(";", paren_end)
        args = text[paren + 1:paren_end]
        out += text[pos:k]
        out += "emu::launch(%s, %s, %s, %s, [&] { %s(%s); })" % (config[0], config[1], config[2], config[3], kernel, args)
        out += text[paren_end + 1:semi + 1]
        pos = semi + 1


def drop_device_only_blocks(text):
    """Remove the `#ifndef MIF_SIMT_EMU` branches (inline PTX the interpreter replaces by emu:: calls) and the
    `#ifdef MIFGPU_PHASE_TRACE` branches (diagnostic build only); the `#else` branch, if any, stays."""
    out, keep, depth_stack = [], True, []
    for line in text.split("\n"):
        stripped = line.strip()
        if stripped.startswith("#ifndef MIF_SIMT_EMU") or stripped.startswith("#ifdef MIFGPU_PHASE_TRACE"):
            depth_stack.append("emu")
            keep = False
            out.append("#if 1  // MIF_SIMT_EMU branch kept by rewrite.py")
            continue
        if stripped.startswith("#if") and depth_stack:
            depth_stack.append("other")
        elif stripped.startswith("#else") and depth_stack and depth_stack[-1] == "emu":
            keep = True
            continue
        elif stripped.startswith("#endif") and depth_stack:
            kind = depth_stack.pop()
            if kind == "emu":
                keep = True
                out.append("#endif")
                continue
        if keep:
            out.append(line)
    return "\n".join(out)


def rewrite(text):
    text = drop_device_only_blocks(text)
    text = rewrite_launches(text)
    text = re.sub(r"extern\s+__shared__\s+([\w:]+(?:\s+[\w:]+)*?)\s+(\w+)\s*\[\s*\]\s*;",
                  r"\1 *\2 = reinterpret_cast<\1 *>(emu::dynamic_smem());", text)
    text = re.sub(r'asm\s+volatile\s*\(\s*"bar\.sync\s+%0,\s*%1;"\s*::\s*"r"\((.*?)\),\s*"r"\((.*?)\)\s*:\s*"memory"\s*\)\s*;',
                  r"emu::named_barrier(\1, \2);", text)
    text = re.sub(r'asm\s+volatile\s*\(\s*"bar\.sync\s+(\d+),\s*(\d+);"\s*:::\s*"memory"\s*\)\s*;',
                  r"emu::named_barrier(\1, \2);", text)
    text = re.sub(r'asm\s+volatile\s*\(\s*"prefetch\.global\.L2\s+\[%0\];"\s*::\s*"l"\((.*?)\)\s*\)\s*;', r"(void)(\1);", text)
    if "asm volatile" in text or "<<<" in text or "extern __shared__" in text:
        raise SystemExit("rewrite.py: a CUDA-only construct was left untranslated")
    return text


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        result = rewrite(f.read())
    with open(dst, "w") as f:
        f.write("// generated by tests/simt_emu/rewrite.py from %s -- do not edit\n" % src)
        f.write(result)
