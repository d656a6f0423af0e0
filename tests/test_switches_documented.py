"""Every experiment switch the library or the Python package reads from the environment (``KOA_*``) has a row in
DESIGN.md section 9b, and every row there is still read somewhere."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read_switches():
    found = set()
    for base, exts in ((os.path.join(ROOT, "oaprogressionmmf_b200"), (".cu", ".cuh", ".h", ".py")),):
        for dirpath, _, files in os.walk(base):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith(exts):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    found |= set(re.findall(r'(?:getenv|env_int|environ\.get)\(\s*"(KOA_[A-Z0-9_]+)"', src))
    return found


def test_switch_table_is_complete_and_current():
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    table = design[design.index("## 9b."):design.index("## 10.")]
    documented = set(re.findall(r"^\| `(KOA_[A-Z0-9_]+)` \|", table, re.M))
    read = _read_switches()
    assert read, "no switches found: the scan is broken"
    assert read - documented == set(), f"read but not documented in DESIGN.md 9b: {sorted(read - documented)}"
    assert documented - read == set(), f"documented but no longer read: {sorted(documented - read)}"
