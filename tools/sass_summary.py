"""Dump per-kernel SASS mnemonic counts of the built extension into profiles/sass/mnemonic_summary.json (runs offline:
cuobjdump reads the in-tree .so).  The mnemonics that prove the Blackwell paths: UTCHMMA / UTCBAR / LDTM (tcgen05 +
TMEM), UTMALDG (TMA), REDG / ATOMG (fused push), SYNCS (mbarrier), PREFETCH."""
import collections
import glob
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "harmony_b200", "ops", "_build", "_hb_native*.so"))[0]
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDGMC", "REDG", "RED", "ATOMG",
        "ATOM", "ATOMS", "SYNCS", "LDG", "STG", "LDS", "STS", "FFMA", "DFMA", "HMMA", "SHFL", "BAR", "CCTL", "MUFU",
        "UCGABAR", "ELECT", "FENCE", "MEMBAR", "LDGSTS", "PREFETCH")
out, cur = {}, None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        out[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        out[cur]["total_instructions"] += 1
        base = m.group(1).split(".")[0]
        if base in KEEP:
            out[cur][base] += 1
summary = {k: dict(sorted(v.items())) for k, v in sorted(out.items()) if k.startswith("_ZN2hb")}
os.makedirs(os.path.join(ROOT, "profiles", "sass"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "sass", "mnemonic_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
print(len(summary), "kernels")
for k, v in summary.items():
    flags = [n for n in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGMC", "UTCBAR", "SYNCS", "REDG", "ATOMG", "MEMBAR") if v.get(n)]
    print(f"{k[:70]:<70} {v.get('total_instructions', 0):>6}  {' '.join(f'{n}={v[n]}' for n in flags)}")

# full listings of the round-2 kernels the verdict asked for (multimem / bulk TMA / flags / stream-K)
WANT = ("ship3_fwd_kernelILi1E", "ship_softmax_kernelILi8ELb0ELb1E", "ship_softmax_kernelILi8ELb1ELb1E",
        "mlr_bwd_push_warp_kernelILi8E", "migrate_bulk_kernel", "migrate_drain_kernelILb0E", "flip_blocks_kernel",
        "pull_rows_lda_sparse_kernel", "dense_gemm_tf32_2cta_kernelILb0ELb0E", "dense_gemm_tf32_sk_kernelILb0ELb0E",
        "mlr_fwd_kernelILi1ELb1E")
cur, buf = None, []
def flush():
    if cur and any(w in cur for w in WANT):
        with open(os.path.join(ROOT, "profiles", "sass", cur[:80] + ".sass"), "w") as f:
            f.write("\n".join(buf) + "\n")
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, buf = m.group(1), [line]
    elif cur:
        buf.append(line)
flush()
