"""SASS evidence for the shipped library (no GPU needed): per-kernel counts of the mnemonics that prove the Blackwell
paths (UTCHMMA / UTCBAR = tcgen05.mma / commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = 1-D bulk TMA, SYNCS = mbarrier,
FFMA2 / FADD2 = packed fp32, MUFU.EX2) and, for the tensor-core kernels, the instruction schedule of the epilogue's
inner body (first LDTM to the following accumulator release) — the evidence VERDICT r01 asked to commit.

    python tools/sass_summary.py r02      -> profiles/r02_sass_summary.json, profiles/r02_sass_<kernel>_epilogue.txt
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "geomloss_b200", "libb200ot.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "MUFU.EX2",
        "MUFU.RSQ", "MUFU.SQRT", "MUFU.LG2", "LDS", "STS", "LDG", "STG", "BAR", "FMNMX", "FMNMX3", "SHFL", "ATOM", "RED"]
WANT = {
    "softmin_partial_D3_big": r"softmin_partial_kernel<b200ot::SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 16, 3, true>, false>",
    "rowsum_softmin_bwd_D3_big": r"rowsum_partial_kernel<b200ot::RowSumCfg<0, 3, 2, 256, 1024, 3, 2>, false>",
    "tc_reduce_conv": r"tc_reduce_kernel<b200ot::TcCfg<128, 16>, 0>",
    "tc_reduce_softmin": r"tc_reduce_kernel<b200ot::TcCfg<128, 16>, 1>",
    # <Cfg, MODE, PT, LDALL, MERGE>: the shipped instantiations (PT = 2, merged hi.[Y_h | Y_l] instruction) and, for
    # comparison, the three-instruction GEMM 2 they replace
    "tc_bwd_conv": r"tc_bwd_kernel<b200ot::TcCfg<128, 8>, 2, 2, false, true>",
    "tc_bwd_softmin": r"tc_bwd_kernel<b200ot::TcCfg<128, 8>, 3, 2, false, true>",
    "tc_bwd_conv_unmerged": r"tc_bwd_kernel<b200ot::TcCfg<128, 8>, 2, 2, false, false>",
    "grid_pass_p2": r"grid_pass_kernel<2, 32>",
    "sinkhorn_iteration_small_D3": r"sinkhorn_iteration_small_kernel<3, 2, 8>",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    chunks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    assert len(chunks) == len(names)
    out = {"library": os.path.relpath(LIB, ROOT), "kernels": {}}
    total = collections.Counter()
    for name, body in zip(names, chunks):
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
        cnt = collections.Counter()
        for op in ops:
            base = op.split(".")[0]
            for k in KEYS:
                if op == k or base == k or op.startswith(k + "."):
                    cnt[k] += 1
        total.update(cnt)
        for short, pat in WANT.items():
            if pat in name:
                regs = re.search(r"REG:(\d+)", body)
                out["kernels"][short] = {"function": name.split("(")[0], "instructions": len(ops),
                                          **{k: cnt[k] for k in KEYS if cnt[k]}}
                if short.startswith("tc_"):
                    lines = [l for l in body.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
                    first = next(i for i, l in enumerate(lines) if "LDTM" in l)
                    stop = next((i for i in range(first + 1, len(lines)) if "SYNCS.ARRIVE" in lines[i]), first + 200)
                    excerpt = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines[max(first - 4, 0):stop + 2]]
                    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_{short}_epilogue.txt"), "w") as f:
                        f.write(f"// {name.split('(')[0]}\n// epilogue body: first LDTM .. accumulator release "
                                f"(cuobjdump -sass {os.path.relpath(LIB, ROOT)})\n" + "\n".join(excerpt) + "\n")
    out["whole_library"] = {k: total[k] for k in KEYS if total[k]}
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
