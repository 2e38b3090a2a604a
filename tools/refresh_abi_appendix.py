"""Rewrites INTEGRATION.md's appendix table from `python tools/abi_table.py` (tests/test_abi.py keeps them in step)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "abi_table.py")], capture_output=True, text=True,
                       check=True).stdout.rstrip("\n")
path = os.path.join(ROOT, "INTEGRATION.md")
doc = open(path).read()
start = doc.index("| function | declared and documented at")
doc = doc[:start] + table + "\n"
open(path, "w").write(doc)
print(table.splitlines()[-1])
