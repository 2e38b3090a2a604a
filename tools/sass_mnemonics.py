"""Per-kernel counts of the SASS mnemonics that show tcgen05 / TMA / tensor-memory use (B200_PROFILING.md: `UTC*MMA` =
tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor, `UTCBAR` = tcgen05.commit; `HMMA` would be the legacy
mma.sync path), read out of the built library with cuobjdump. Runs without a GPU.
    python tools/sass_mnemonics.py > profiles/r1_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|HMMA|HGMMA|REDG?|ATOMG)\b")


def main():
    so = os.path.join(ROOT, "scaledreamer_b200", "libsdb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = PAT.search(line)
        if m and fn:
            cnt[fn][m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass scaledreamer_b200/libsdb200.so (sm_100a), mnemonic counts per kernel")
    for name, c in sorted(zip(names, cnt.values())):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        print(f"{short:48s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    legacy = [n for n, c in zip(names, cnt.values()) if c.get("HMMA") or c.get("HGMMA")]
    print(f"# kernels using the legacy mma.sync tensor path (HMMA): {len(legacy)}")


if __name__ == "__main__":
    main()
