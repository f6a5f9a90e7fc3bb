"""Compare two builds of libttmpc.so kernel by kernel at the SASS instruction level (development aid):
    python tools/sass_diff.py old.so new.so
Prints SAME / DIFF per kernel; used to show that a source change that is compiled out (experiment macros off) leaves the
shipped, GPU-validated code untouched."""
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_8_ttmpc_cu_[0-9a-f]+", "", m.group(1))  # per-build hash of the anonymous namespace
            out[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur is not None:
            out[cur].append(m.group(1).strip())
    return out


if __name__ == "__main__":
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    bad = 0
    for k in sorted(set(a) | set(b)):
        same = a.get(k) == b.get(k)
        bad += not same
        print(("SAME " if same else "DIFF "), len(a.get(k, [])), len(b.get(k, [])), k[-70:])
    print(f"{bad} of {len(set(a) | set(b))} kernels differ")
    sys.exit(1 if bad else 0)
