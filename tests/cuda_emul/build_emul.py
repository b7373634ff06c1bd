"""tests/cuda_emul/build_emul.py — TEST INFRASTRUCTURE.

Rewrites the product's CUDA sources (prefhetch_b200/csrc) into plain C++ and compiles them with g++ against
tests/cuda_emul/include/cuda_runtime.h (a small CPU interpreter of the CUDA execution model), producing
`libprefhetch_b200_emul.so` in a scratch directory.  The `-m gpu` parity tests can then be dry-run on a machine
without a GPU (tests/test_cuda_emulated.py): same kernels, same host code, same C ABI, executed block by block on
host threads.  Nothing in the package loads this library on its own; it is selected with PF_LIB by the test.

The rewrite is textual and deliberately narrow — it fails loudly when the sources change in a way it does not know:
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  pf_emul::launch(grid, block, smem, [&]{ kernel(args); })
  * `extern __shared__ [__align__(n)] T name[];`   ->  T *name = (T *)pf_emul::dyn_smem();
  * every inline-PTX statement of the sources      ->  its C equivalent (table ASM below; an unknown `asm` aborts)
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
CSRC = ROOT / "prefhetch_b200" / "csrc"
HERE = Path(__file__).resolve().parent

# inline PTX of the product -> C.  Keys are matched after whitespace normalisation of the whole asm statement.
ASM = {
    # pf_common.cuh: cache-hinted 128-bit loads / stores, L2 policies (no cache hierarchy here)
    'asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));': "r = *p;",
    'asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));': "pol = 1;",
    'asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));': "pol = 2;",
    'asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));': "pol = 3;",
    'asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol));':
        "(void)pol; r = *p;",
    'asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(pol));':
        "(void)pol; *p = v;",
    'asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y));': "*p = v;",
    # pf_mac.cuh: acc += a*b as one 32x32->64 multiply-add (carry chain spelled in PTX)
    'asm("{\\n\\t" ".reg .u32 l0, l1;\\n\\t" "mov.b64 {l0, l1}, %0;\\n\\t" "mad.lo.cc.u32 l0, %1, %2, l0;\\n\\t" "madc.hi.u32 l1, %1, %2, l1;\\n\\t" '
    '"mov.b64 %0, {l0, l1};\\n\\t" "}" : "+l"(acc) : "r"(a), "r"(b));': "acc += (u64)a * (u64)b;",
    # 128-bit accumulate of a 64x64 product
    'asm("mad.lo.cc.u64 %0, %2, %3, %0;\\n\\t" "madc.hi.u64 %1, %2, %3, %1;" : "+l"(a.lo), "+l"(a.hi) : "l"(x), "l"(y));':
        "{ unsigned __int128 s_ = ((unsigned __int128)a.hi << 64 | a.lo) + (unsigned __int128)x * y; a.lo = (u64)s_; a.hi = (u64)(s_ >> 64); }",
    'asm("mov.b64 {%0, %1}, %2;" : "=r"(o.x0), "=r"(o.x1) : "l"(w));': "o.x0 = (u32)w; o.x1 = (u32)(w >> 32);",
    # pf_plain.cuh: nanosecond timer of the bounded flag wait
    'asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));': "t0 = pf_emul::globaltimer();",
    'asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));': "t1 = pf_emul::globaltimer();",
}


def _norm(s: str) -> str:
    return re.sub(r"\s+", " ", s).strip()


ASM_N = {_norm(k): v for k, v in ASM.items()}


def _match_paren(text: str, i: int, open_c="(", close_c=")") -> int:
    """index just past the bracket that closes text[i] (string literals skipped)"""
    assert text[i] == open_c, (text[i - 20:i + 20])
    depth, j = 0, i
    while j < len(text):
        c = text[j]
        if c == '"':
            j += 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
        elif c == "'":
            j += 1
            while text[j] != "'":
                j += 2 if text[j] == "\\" else 1
        elif c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return j + 1
        j += 1
    raise ValueError("unbalanced bracket")


def rewrite_asm(text: str, name: str) -> str:
    out, i = [], 0
    for m in re.finditer(r"\basm\s*(volatile\s*)?\(", text):
        if m.start() < i:
            continue
        # skip mentions inside comments
        line_start = text.rfind("\n", 0, m.start()) + 1
        if "//" in text[line_start:m.start()]:
            continue
        end = _match_paren(text, m.end() - 1)
        if text[end] != ";":
            raise SystemExit(f"{name}: asm statement not followed by ';'")
        stmt = _norm(text[m.start():end + 1])
        # adjacent string literals may be split differently: compare with the literals' seams normalised too
        key = stmt
        if key not in ASM_N:
            raise SystemExit(f"{name}: inline PTX the emulation does not know:\n  {stmt}\n(add its C equivalent to ASM in {__file__})")
        out.append(text[i:m.start()])
        out.append(ASM_N[key])
        out.append("\n" * text[m.start():end + 1].count("\n"))
        i = end + 1
    out.append(text[i:])
    return "".join(out)


def _split_top(s: str):
    parts, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur))
    return [p.strip() for p in parts]


def rewrite_launches(text: str, name: str) -> str:
    out, i, n = [], 0, 0
    while True:
        k = text.find("<<<", i)
        if k < 0:
            break
        line_start = text.rfind("\n", 0, k) + 1
        if "//" in text[line_start:k]:
            out.append(text[i:k + 3])
            i = k + 3
            continue
        # kernel expression: identifier, optionally with template arguments, directly before <<<
        j = k
        if text[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        s = j
        while s > 0 and (text[s - 1].isalnum() or text[s - 1] in "_:"):
            s -= 1
        kern = text[s:k]
        e = text.find(">>>", k)
        cfg = _split_top(text[k + 3:e].replace("\\\n", " "))
        if not 2 <= len(cfg) <= 4:
            raise SystemExit(f"{name}: launch configuration with {len(cfg)} arguments: {text[k:e + 3]}")
        a = e + 3
        while text[a] in " \t\n\\":
            a += 1
        aend = _match_paren(text, a)
        args = text[a + 1:aend - 1].replace("\\\n", " ")
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(text[i:s])
        out.append(f"pf_emul::launch(pf_emul::mkdim({cfg[0]}), pf_emul::mkdim({cfg[1]}), (size_t)({smem}), [&]() {{ {kern}({args}); }})")
        i = aend
        n += 1
    out.append(text[i:])
    return "".join(out)


def rewrite_dyn_smem(text: str) -> str:
    pat = re.compile(r"extern\s+__shared__\s+(?:__align__\(\s*\d+\s*\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;")
    return pat.sub(lambda m: f"{m.group(1)} *{m.group(2)} = reinterpret_cast<{m.group(1)} *>(pf_emul::dyn_smem());", text)


def rewrite(text: str, name: str) -> str:
    text = rewrite_asm(text, name)
    text = rewrite_launches(text, name)
    text = rewrite_dyn_smem(text)
    if "<<<" in re.sub(r"//[^\n]*", "", text) or re.search(r"\bextern\s+__shared__", text):
        raise SystemExit(f"{name}: a launch or a dynamic shared-memory declaration survived the rewrite")
    return text


def asan_runtime() -> str:
    """libasan.so to LD_PRELOAD into the python that loads an --asan build"""
    return subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()


def build(out_dir: Path, opt: str = "-O2", verbose: bool = False, asan: bool = False) -> Path:
    out_dir = Path(out_dir)
    gen = out_dir / "gen" / "prefhetch_b200" / "csrc"
    gen.mkdir(parents=True, exist_ok=True)
    inc = out_dir / "gen" / "include"
    inc.mkdir(parents=True, exist_ok=True)
    (inc / "prefhetch_b200.h").write_text((ROOT / "include" / "prefhetch_b200.h").read_text())
    units = []
    for p in sorted(CSRC.iterdir()):
        if p.suffix not in (".cu", ".cuh", ".h"):
            continue
        dst = gen / (p.name + ".cpp" if p.suffix == ".cu" else p.name)
        dst.write_text(rewrite(p.read_text(), p.name))
        if p.suffix == ".cu":
            units.append(dst)
    so = out_dir / ("libprefhetch_b200_emul_asan.so" if asan else "libprefhetch_b200_emul.so")
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []
    if asan:
        opt = "-O1"
    cmd = ["/usr/bin/g++", "-std=c++17", opt, *san, "-g1", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-pthread",
           "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes",
           "-I" + str(HERE / "include"), "-o", str(so), *[str(u) for u in units], "-lz", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (out_dir / "build_emul.log").write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode:
        sys.stderr.write((r.stdout + r.stderr)[-20000:])
    if r.returncode:
        raise RuntimeError("g++ failed on the rewritten CUDA sources (see build_emul.log)")
    return so


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    d = Path(argv[0] if argv else os.environ.get("PF_EMUL_DIR", "/tmp/pf_cuda_emul"))
    print(build(d, verbose=True, asan="--asan" in sys.argv))
