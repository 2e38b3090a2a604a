"""Prints the INTEGRATION.md appendix: every function the C ABI exports (include/*.h), the header that documents it
(with the reference call site it replaces) and the host-side module that binds it.
    python tools/abi_table.py > /tmp/abi.md"""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    hosts = {p: open(p).read() for p in glob.glob(os.path.join(ROOT, "scaledreamer_b200", "*.py"))}
    rows = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        src = open(h).read()
        # "<type> sdb_name(" at the start of a declaration
        for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(sdb_\w+)\s*\(", src, flags=re.M):
            name = m.group(1)
            line = src[:m.start()].count("\n") + 1
            users = sorted(os.path.basename(p) for p, t in hosts.items() if re.search(r"\b%s\b" % name, t))
            users = [u for u in users if u != "lib.py"] or [u + " (signature / loader)" for u in users]
            rows.append((name, f"{os.path.basename(h)}:{line}", ", ".join(users) or "— (exported for other hosts / diagnostics)"))
    print("| function | declared and documented at | bound by (scaledreamer_b200/) |")
    print("|---|---|---|")
    for r in rows:
        print("| `%s` | `include/%s` | %s |" % r)
    print(f"\n{len(rows)} functions.")


if __name__ == "__main__":
    main()
