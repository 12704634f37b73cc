"""Per-kernel SASS opcode counts of the built library (the evidence that the hot kernels are tcgen05 / TMEM / TMA code).
usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pytorch_openpose_b200", "libopenpose_b200.so")
KEEP = re.compile(r"^(UTCHMMA|LDTM|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|SYNCS|ACQBULK|PREEXIT|HMMA|DADD|DMUL|DFMA|FFMA2?|UCGABAR)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = sass
    per = collections.OrderedDict()
    cur = None
    for line in names.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("opb::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
            cur = re.sub(r"\(.*$", "", cur).replace("opb::", "")
            per[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            if KEEP.match(op):
                # keep the variant suffixes that matter (.2CTA, .MULTICAST, dimensions), drop operand-size noise
                parts = op.split(".")
                key = ".".join(p for p in parts if p in (parts[0], "2CTA", "1CTA", "2D", "4D", "5D", "MULTICAST", "x32", "x16"))
                per[cur][key] += 1
    print("# per-kernel SASS opcode counts of pytorch_openpose_b200/libopenpose_b200.so (cuobjdump -sass, sm_100a), round 2 final tree")
    print("# tcgen05.mma -> UTCHMMA[.2CTA]; tcgen05.ld -> LDTM; TMA loads -> UTMALDG.*; tcgen05.commit -> UTCBAR*; tcgen05.alloc -> UTCATOMSWS; mbarrier -> SYNCS;")
    print("# griddepcontrol.wait -> ACQBULK; griddepcontrol.launch_dependents -> PREEXIT (programmatic dependent launch)")
    print()
    for name, c in per.items():
        if c:
            print("%-60s %s" % (name, "  ".join("%s=%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
