"""SASS evidence for the Blackwell-native paths: per-kernel counts of the tell-tale mnemonics plus verbatim instruction
excerpts (one kernel per family).  Runs without a GPU (cuobjdump on the in-tree .so).
    python benchmarks/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "poseidon_b200/_ext/poseidon_b200_C.so"
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTCBAR[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCALLOC|UTCATOMSWS[.\w]*|"
                 r"LDGMC[.\w]*|STGMC[.\w]*|REDGMC[.\w]*|LDGSTS[.\w]*|ARRIVES[.\w]*|SYNCS[.\w]*|UCGABAR[.\w]*|ELECT|HMMA[.\w]*)")
text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        funcs[cur] = []
    elif cur is not None and "/*" in line and ";" in line:
        funcs[cur].append(line)
print(f"# cuobjdump -sass {SO} (sm_100a): {len(funcs)} kernels")
print("# PTX -> SASS: tcgen05.mma = UTCHMMA(.2CTA), tcgen05.commit = UTCBAR(.2CTA)(.MULTICAST), tcgen05.ld = LDTM, TMA load = UTMALDG"
      "(.4D.IM2COL)(.2CTA)(.MULTICAST), bulk store = UBLKCP, multimem.ld_reduce = LDGMC.E.ADD..., "
      "multimem.st = STG.E.128.STRONG.SYS to the multicast address (no mnemonic of its own), "
      "cp.async = LDGSTS, HMMA = legacy mma.sync (none expected)")
print("\n## 1. mnemonic counts per kernel\n")
for name, lines in funcs.items():
    c = collections.Counter()
    for ln in lines:
        for mm in PAT.findall(ln):
            c[mm.split(".64")[0]] += 1
    if c:
        short = re.sub(r"\(.*", "", name.replace("psd::", "").replace("void ", ""))
        print(f"{short[:78]:78s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))


def excerpt(match, want, title, ctx=1, limit=14):
    for name, lines in funcs.items():
        if match(name):
            hits = [i for i, ln in enumerate(lines) if re.search(want, ln)]
            if not hits:
                continue
            print(f"\n### {title}\n### {name[:150]}")
            shown = set()
            for i in hits[:limit]:
                for j in range(max(0, i - ctx), min(len(lines), i + ctx + 1)):
                    if j not in shown:
                        shown.add(j)
                        print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", lines[j]).rstrip())
            return
    print(f"\n### {title}: NOT FOUND")


print("\n## 2. instruction excerpts\n")
excerpt(lambda n: "umma_gemm_kernel<256, false, false, 0, 0, 2>" in n, r"UTCHMMA|UTCBAR|UTMALDG|LDTM|UTCALLOC",
        "paired-CTA GEMM (cta_group::2): 2CTA MMA, multicast commit, 2CTA TMA loads, TMEM loads")
excerpt(lambda n: "umma_gemm_kernel<192, false, false, 0, 3, 1>" in n, r"UTMALDG|UTCHMMA", "conv fprop: im2col-mode TMA + tcgen05.mma")
excerpt(lambda n: "umma_gemm_kernel<192, false, false, 0, 3, 2>" in n, r"UTMALDG", "conv fprop, paired: im2col-mode TMA in its 2CTA form")
excerpt(lambda n: "umma_gemm_kernel<128, false, false, 0, 3, 1>" in n, r"UTMALDG.*MULTICAST|UTCBAR.*MULTICAST",
        "conv fprop, cluster multicast of the im2col operand")
excerpt(lambda n: "umma_gemm_kernel<256, true, true, 1, 0, 2>" in n, r"UBLKCP|UTCHMMA", "fp32 epilogue: bulk-async (TMA) row stores")
excerpt(lambda n: "umma_gemm_kernel<256, true, true, 2, 0, 2>" in n, r"UTCHMMA|LDTM|LDG\.E\.CONSTANT|STG", "SFB reconstruct with the optimizer epilogue", ctx=0, limit=10)
excerpt(lambda n: "allreduce_sgd_multi_kernel" in n, r"LDGMC|STGMC|REDGMC|MULTIMEM|LD\.E\..*SYS|ST\.E\..*SYS|MEMBAR",
        "fused all-reduce + SGD: multimem.ld_reduce / multimem.st (NVLS) and system-scope flags", ctx=0, limit=16)
excerpt(lambda n: "ssp_fold_kernel" in n, r"ST\.E\..*SYS|LD\.E\..*SYS|MEMBAR|ATOMG", "SSP fold: peer loads + system-scope consumed flags", ctx=0, limit=8)
excerpt(lambda n: "peer_push_kernel" in n, r"STG\.E[.\w]*STRONG\.SYS|STG\.E\.128", "SFB factor publish: stores to the multicast (or peer) address", ctx=0, limit=6)
