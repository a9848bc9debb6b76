"""Static SASS mnemonic counts per kernel of the built library (no GPU needed):
python tools/sass_counts.py > profiles/r2_sass_tensor_tma_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rohm_b200", "librohm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCBAR", "UTCATOMSWS", "HMMA", "MUFU.EX2")
KEEP = re.compile(r"gemm_tile_kernelILi(128|96|64|32)ELi3ELi\dELi1E|attention_tc|ddpm_step_philox|gn_mish|sum_split|guide_backward|"
                  r"skin_kernel|fk_full|projection_guidance")
print("SASS mnemonic counts per kernel (cuobjdump -sass rohm_b200/librohm_b200.so, sm_100a; static instruction counts).")
print("UTCHMMA = tcgen05.mma kind::f16 / kind::tf32, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,")
print("UBLKCP = bulk copy, SYNCS = mbarrier operations, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc / dealloc, HMMA = legacy mma.sync.")
print("gemm_tile_kernel<BLOCK_N, PASSES, EPI, KIND>: KIND 1 = fp16 pairs (the default mode), 0 = TF32.\n")
name, counts, total = None, collections.Counter(), 0
def flush():
    if name and KEEP.search(name):
        short = KEEP.search(name).group(0)
        m2 = re.search(r"\d+([a-z_]+kernel)", name)
        if not short.startswith("gemm_tile") and m2:
            short = m2.group(1)
        print(f"{short[:44]:44s} {total:6d} instr  " + "  ".join(f"{k} {v}" for k, v in sorted(counts.items())))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        total += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                counts[w] += 1
flush()
