"""Per-kernel SASS mnemonic counts of libpfpp_sm100.so (cuobjdump -sass): the evidence that the tensor-core kernels are
tcgen05 + TMEM + TMA (UTCHMMA / LDTM / STTM / UTMALDG / UTMASTG) and that no legacy mma.sync (HMMA) path exists.
    python profiles/sass_table.py [path/to/libpfpp_sm100.so] > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

MN = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "LDGSTS", "REDUX", "HMMA", "FFMA", "MUFU.EX2"]


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"^void |\(anonymous namespace\)::|_GLOBAL__N_\w+?_cu_[0-9a-f]+::", "", fn)
            fn = re.sub(r"\((?:CUtensorMap_st|float|int|void|long|unsigned|__nv_bfloat16|\w+ const\*).*", "", fn)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[fn]
        if op.startswith("UTCHMMA"):
            c["UTCHMMA"] += 1
            if ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
        elif op.startswith("HMMA"):
            c["HMMA"] += 1
        elif op.startswith("MUFU.EX2"):
            c["MUFU.EX2"] += 1
        else:
            for k in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "LDGSTS", "REDUX", "FFMA"):
                if op.startswith(k):
                    c[k] += 1
    print(f"{'kernel':84s} " + " ".join(f"{m:>12s}" for m in MN))
    tot = collections.Counter()
    for fn, c in sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], kv[0])):
        print(f"{fn[:84]:84s} " + " ".join(f"{c[m]:12d}" for m in MN))
        tot.update(c)
    print(f"{'TOTAL (%d kernels)' % len(counts):84s} " + " ".join(f"{tot[m]:12d}" for m in MN))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "puzzlefusion-plusplus_b200", "libpfpp_sm100.so"))
