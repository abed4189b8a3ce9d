"""SASS opcode histogram of the shipped library (cuobjdump -sass): per kernel, the mnemonics that prove which
hardware path it uses — UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st on TMEM), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (TMA tensor load / store), SYNCS (mbarrier), HMMA (legacy mma.sync),
LDGSTS (cp.async), REDG / ATOMG, plus the total instruction count.
usage: python profiles/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genie-tts_b200", "csrc", "libgenie_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA",
        "LDGSTS", "REDG", "ATOMG", "FFMA", "LDG", "STG", "LDS", "STS", "BAR", "ACQBULK", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                          text=True).stdout.splitlines()
    names = iter(filt)
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            cur = re.sub(r"\(genie::.*|\((?:int|float|long|void|unsigned|__half|genie)[^<]*$", "", cur)
            cur = cur.replace("genie::(anonymous namespace)::", "").replace("void ", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels (sm_100a)")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# whole library: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
    print(f"{'kernel':70s} {'instr':>7s}  key opcodes")
    for name, c in sorted(per.items(), key=lambda kv: -kv[1]['_total']):
        keys = "  ".join(f"{k}={c[k]}" for k in KEYS[:14] if c[k])
        print(f"{name[:70]:70s} {c['_total']:7d}  {keys}")


if __name__ == "__main__":
    main()
