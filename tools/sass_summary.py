"""Per-kernel SASS evidence for the Blackwell-native claims: counts of the tcgen05 / TMEM / TMA / packed-fp32
mnemonics in every kernel of libmvfusion.so.  Regenerate with:  python tools/sass_summary.py > profiles/sass_summary.txt
(needs cuobjdump from the CUDA toolkit and c++filt; runs on the build container, no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mulit_view_object_detection_b200", "libmvfusion.so")
# mnemonic prefix -> what it proves (/opt/skills/guides/B200_PROFILING.md)
WATCH = [("UTCHMMA", "tcgen05.mma kind::f16 (5th-gen tensor core, smem descriptors)"),
         ("UTCQMMA", "tcgen05.mma kind::tf32 / fp8"),
         ("UTCBAR", "tcgen05.commit -> mbarrier"),
         ("LDTM", "tcgen05.ld (TMEM -> registers)"),
         ("STTM", "tcgen05.st (registers -> TMEM)"),
         ("UTMALDG", "TMA tensor load (cp.async.bulk.tensor global->shared)"),
         ("UTMASTG", "TMA tensor store (shared->global)"),
         ("UBLKCP", "cp.async.bulk (non-tensor bulk copy)"),
         ("SYNCS", "mbarrier arrive / try_wait"),
         ("FFMA2", "packed fp32x2 FMA"),
         ("FFMA", "fp32 FMA (scalar; includes FFMA2)"),
         ("LDG", "global load"), ("STG", "global store"), ("LDS", "shared load"), ("STS", "shared store"),
         ("SHFL", "warp shuffle"), ("VOTE", "warp ballot"), ("REDUX", "warp reduce")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur]["_total"] += 1
            op = m.group(1)
            for key, _ in WATCH:
                if op.startswith(key):
                    kernels[cur][key] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("SASS summary of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("legend:")
    for key, what in WATCH:
        print("  %-8s %s" % (key, what))
    tot = collections.Counter()
    rows = []
    for mangled, name in zip(names, dem):
        c = kernels[mangled]
        short = re.sub(r"\(.*", "", name).replace("mvf::", "")
        rows.append((short, c))
        tot.update(c)
    keys = [k for k, _ in WATCH if tot[k]]
    print()
    print("%-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in keys))
    for short, c in sorted(rows, key=lambda r: -r[1]["_total"]):
        print("%-78s %7d " % (short[:78], c["_total"]) + " ".join("%7d" % c[k] for k in keys))
    print("%-78s %7d " % ("TOTAL (%d kernels)" % len(rows), tot["_total"]) + " ".join("%7d" % tot[k] for k in keys))


if __name__ == "__main__":
    sys.exit(main())
