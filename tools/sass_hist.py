"""Developer tool: per-kernel SASS opcode histogram of pearl_b200/libpearlb200.so (cuobjdump -sass) -> profiles/sass_opcodes.txt.
The columns are the mnemonics that show which hardware path a kernel uses (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, LDGSTS = cp.async."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "BAR", "MUFU"]


def main():
    so = os.path.join(ROOT, "pearl_b200", "libpearlb200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    out, name, hist = [], None, None

    def flush():
        if name is not None:
            out.append((name, hist))
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            mangled = m.group(1)
            dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
            name = dem.replace("(anonymous namespace)::", "").replace("prl::", "")
            name = re.sub(r"^void ", "", name)
            name = re.sub(r"\((?!int\)|bool\)).*", "", name).replace("(int)", "").replace("(bool)", "")
            hist = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and hist is not None:
            hist[m.group(1).split(".")[0]] += 1
            hist["total"] += 1
    flush()
    w = max(len(n) for n, _ in out) + 2
    lines = ["# SASS opcode histogram of pearl_b200/libpearlb200.so (cuobjdump -sass), round 2, per kernel (tools/sass_hist.py).",
             "# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA),",
             "# SYNCS = mbarrier operations, LDGSTS = cp.async, HMMA = legacy mma.sync (none)",
             "kernel".ljust(w) + "".join(c.rjust(8) for c in COLS + ["total"])]
    seen = set()
    for n, h in out:
        key = (n, tuple(sorted(h.items())))      # templates instantiated in several translation units
        if key in seen:
            continue
        seen.add(key)
        lines.append(n.ljust(w) + "".join(str(h.get(c, 0)).rjust(8) for c in COLS + ["total"]))
    path = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, len(out), "kernels")


if __name__ == "__main__":
    main()
