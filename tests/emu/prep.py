"""TEST INFRASTRUCTURE -- rewrites the product's CUDA sources (medical-transformer_b200/csrc) into plain C++ for the
CUDA-on-CPU shim (tests/emu/include/cuda_runtime.h) WITHOUT touching them:
  * `kernel<...><<<grid, block, smem, stream>>>(args);`  ->  `emu::launch(grid, block, smem, [&]() { kernel<...>(args); });`
  * `extern __shared__ __align__(16) float NAME[];`       ->  a pointer to the emulator's dynamic shared memory
  * the inline-PTX helpers (ex2, lg2, the TMA bulk-copy issue / wait) get portable bodies
  * the relative include of include/medt_axial.h becomes absolute.
Everything else (indexing, barriers, shuffles, atomics, finalisers, the C ABI entry points) is compiled as written."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "medical-transformer_b200", "csrc")

BODIES = {
    r"float\s+ex2\s*\(\s*float\s+x\s*\)": "return exp2f(x);",
    r"float\s+lg2\s*\(\s*float\s+x\s*\)": "return log2f(x);",
    r"void\s+tma_strips_issue\s*\(\s*uint64_t\s*\*\s*bar\s*,\s*float\s*\*\s*dst\s*,\s*const\s+float\s*\*\s*src\s*,\s*uint32_t\s+bytes\s*\)":
        "(void)bar; memcpy(dst, src, bytes);",
    r"void\s+tma_strips_wait\s*\(\s*uint64_t\s*\*\s*bar\s*\)": "(void)bar;",
    r"void\s+phase_stamp\s*\(\s*const\s+Dims\s*&\s*m\s*,\s*int\s+ph\s*\)": "(void)m; (void)ph;",
    r"void\s+cp_async4\s*\(\s*float\s*\*\s*dst_smem\s*,\s*const\s+float\s*\*\s*src\s*\)": "*dst_smem = *src;",
    r"void\s+cp_async16\s*\(\s*float\s*\*\s*dst_smem\s*,\s*const\s+float\s*\*\s*src\s*\)": "memcpy(dst_smem, src, 16);",
    r"void\s+cp_async_wait_all\s*\(\s*\)": "",
    # grid barrier of the cooperative kernels: CTA barrier, one thread parks at the scheduler's grid barrier, CTA barrier
    r"void\s+grid_barrier\s*\(\s*unsigned\s*\*\s*bar\s*,\s*unsigned\s+nctas\s*,\s*unsigned\s*&\s*epoch\s*\)":
        "(void)bar; (void)nctas; ++epoch; __syncthreads(); if (threadIdx.x == 0) emu::barrier_grid(); __syncthreads();",
}


def _match_forward(s, i, open_ch, close_ch):
    """s[i] == open_ch; returns the index of the matching close_ch."""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(s):
    while True:
        k = s.find("<<<")
        if k < 0:
            return s
        # kernel expression: identifier, optionally followed by a balanced <...> template argument list
        j = k
        while s[j - 1].isspace():
            j -= 1
        if s[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if s[j] == ">":
                    depth += 1
                elif s[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while s[j - 1].isalnum() or s[j - 1] in "_:":
            j -= 1
        kernel = s[j:k].strip()
        e = s.index(">>>", k)
        cfg = _split_top(s[k + 3:e])
        p0 = s.index("(", e)
        p1 = _match_forward(s, p0, "(", ")")
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        repl = f"emu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&]() {{ {kernel}{s[p0:p1 + 1]}; }})"
        s = s[:j] + repl + s[p1 + 1:]


def rewrite_bodies(s):
    for pat, body in BODIES.items():
        m = re.search(pat + r"\s*\{", s)
        if not m:
            continue
        b0 = m.end() - 1
        b1 = _match_forward(s, b0, "{", "}")
        s = s[:b0] + "{ " + body + " }" + s[b1 + 1:]
    return s


def rewrite_static_shared(text):
    """`__shared__ [__align__(n)] T a[N], b[M][K];`  ->  references to per-CTA heap blocks from emu::cta_shared()
    (fresh + poisoned for every CTA; each array its own allocation, so AddressSanitizer sees overruns and
    ThreadSanitizer starts every CTA with clean race state)."""
    out = []
    for line in text.split("\n"):
        m = re.match(r"^(\s*)__shared__\s+(?:__align__\(\d+\)\s+)?(.*);\s*$", line)
        if not m or "extern" in line:
            out.append(line)
            continue
        pieces = _split_top(m.group(2))
        first = re.match(r"^(\w+)\s+(\w+)((?:\[[^\]]*\])*)$", pieces[0])
        assert first, line
        typ = first.group(1)
        decls = [(first.group(2), first.group(3))]
        for pc in pieces[1:]:
            mm = re.match(r"^(\w+)((?:\[[^\]]*\])*)$", pc)
            assert mm, line
            decls.append((mm.group(1), mm.group(2)))
        stm = []
        for name, dims in decls:
            stm.append(f"typedef {typ} emu_t_{name}{dims}; static char emu_k_{name}; "
                       f"emu_t_{name}& {name} = *reinterpret_cast<emu_t_{name}*>(emu::cta_shared(&emu_k_{name}, sizeof(emu_t_{name})));")
        out.append(m.group(1) + " ".join(stm))
    return "\n".join(out)


def convert(text):
    text = text.replace('#include "../../include/medt_axial.h"', f'#include "{os.path.join(ROOT, "include", "medt_axial.h")}"')
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?float\s+(\w+)\[\];",
                  r"float* \1 = reinterpret_cast<float*>(emu::g_dyn_smem);", text)
    text = rewrite_static_shared(text)
    text = rewrite_bodies(text)
    text = rewrite_launches(text)
    assert "<<<" not in text and "asm volatile" not in text.split("__CUDA_ARCH__")[0], "unconverted CUDA construct"
    return text


# tcgen05 / TMEM kernels have no CPU emulation: the emulated library reports "shape not eligible" and the FMA kernels run instead
TENSOR_CORE_STUBS = {"conv_tc.cu": '''#include "conv_tc.cuh"
static int g_mode = 0;
int tc_conv_fwd_or_dgrad(const TcConv&, const float*, const float*, float*, cudaStream_t) { return 0; }
int tc_conv_wgrad(const TcConv&, const float*, const float*, float*, cudaStream_t) { return 0; }
int tc_conv_mode() { return g_mode; }
void tc_conv_set_mode(int m) { g_mode = m; }
''',
                     # thread-block clusters / distributed shared memory: not emulated, the one-CTA-per-channel kernels run instead
                     "bn_cluster.cu": '''#include "bn_cluster.cuh"
int bn_cluster_forward(const BnActArgs&, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float*,
                       cudaStream_t) { return 0; }
int bn_cluster_backward(const BnActArgs&, const float*, const float*, const float*, const float*, const float*, const float*, float*,
                        float*, float*, float*, cudaStream_t) { return 0; }
'''}


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    made = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            dst = os.path.join(out_dir, f[:-3] + ".cpp" if f.endswith(".cu") else f)
            txt = TENSOR_CORE_STUBS[f] if f in TENSOR_CORE_STUBS else convert(open(os.path.join(CSRC, f)).read())
            if not os.path.exists(dst) or open(dst).read() != txt:
                open(dst, "w").write(txt)
            made.append(dst)
    return made


if __name__ == "__main__":
    print("\n".join(main(sys.argv[1])))
