"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md), from the built library:
python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "autoregressive_diffusion_b200", "liboniris_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", ".2CTA", "tmemA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "FFMA2", "FADD2", "MUFU.EX2", "HMMA"]
print("# cuobjdump -sass autoregressive_diffusion_b200/liboniris_b200.so  (sm_100a): tensor-core / TMEM / TMA mnemonics per kernel")
print("# UTCHMMA = tcgen05.mma (bf16), .2CTA = cta_group::2, tmemA = UTCHMMA whose A operand is read from tensor memory (TS mode);")
print("# LDTM/STTM = tcgen05.ld/st; UTMALDG = TMA tensor load; UBLKCP = cp.async.bulk (1-D); UTCBAR = tcgen05.commit;")
print("# FFMA2/FADD2 = packed fp32 pairs (fma/add.f32x2); HMMA (legacy mma.sync) must be absent")
print(f"{'kernel':70s} " + " ".join(f"{c:>8s}" for c in cols))
chunks = re.split(r"Function : \S+", sass)[1:]
for name, body in zip(names, chunks):
    c = collections.Counter()
    for line in body.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if not m:
            continue
        op, rest = m.group(1), m.group(2)
        if op.startswith("UTCHMMA"):
            c["UTCHMMA"] += 1
            c[".2CTA"] += ".2CTA" in op
            c["tmemA"] += rest.strip().startswith("tmem[")
        for k in ("LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "FFMA2", "FADD2", "HMMA"):
            c[k] += op.startswith(k)
        c["MUFU.EX2"] += op == "MUFU.EX2"
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"{short:70s} " + " ".join(f"{c[k]:8d}" for k in cols))
