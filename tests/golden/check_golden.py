"""check_golden.py -- verify the golden fixtures of this directory against the reference's own log files.

The fixtures hold numbers transcribed from the golden logs under <reference>/tests/*.output; every entry names the
line of the log it stands on.  This script re-reads the logs (only where the reference tree is present, i.e. in the
development container) and checks that every recorded value is printed on the cited line, so that a transcription
error cannot survive.  The tests themselves never read the reference tree: they use the committed JSON files.

    python tests/golden/check_golden.py [/root/reference]

Exit code 0 and one line per fixture if everything matches; tests/test_oracle_pins.py runs it when the tree exists.
"""
import glob
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _tokens(value):
    """the strings a recorded value must appear as on its line"""
    if isinstance(value, (list, tuple)):
        return [t for v in value for t in _tokens(v)]
    if isinstance(value, bool):
        return []
    if isinstance(value, int):
        return [str(value)]
    if isinstance(value, float):
        return [f"{value:g}"]
    if isinstance(value, str):
        return [value]
    return []


def _has(line, token):
    # the token must stand alone: not as a part of a longer number
    return re.search(r"(?<![\w.+-])" + re.escape(token) + r"(?![\w]|\.\d)", line) is not None


def check_entry(name, entry, lines, problems):
    if not isinstance(entry, dict):
        return 0
    n = 0
    if "line" in entry:
        text = lines[entry["line"] - 1]
        side = {k[: -len("_line")] for k in entry if k.endswith("_line")}
        for key, value in entry.items():
            if key == "line" or key.startswith("_") or key.endswith("_line") or key in side:
                continue
            for tok in _tokens(value):
                n += 1
                if not _has(text, tok):
                    problems.append(f"{name}.{key} = {tok!r} not on line {entry['line']}: {text.strip()!r}")
        for key in side:                      # values with their own line, e.g. "level_set" / "level_set_line"
            text2 = lines[entry[key + "_line"] - 1]
            for tok in _tokens(entry[key]):
                n += 1
                if not _has(text2, tok):
                    problems.append(f"{name}.{key} = {tok!r} not on line {entry[key + '_line']}: {text2.strip()!r}")
    else:
        for key, value in entry.items():
            if not key.startswith("_") and key != "parameters":
                n += check_entry(f"{name}.{key}", value, lines, problems)
    return n


def check_fixture(path, reference_root):
    with open(path) as f:
        gold = json.load(f)
    m = re.search(r"tests/([\w.]+\.output)", gold["_source"])
    log = os.path.join(reference_root, "tests", m.group(1))
    with open(log) as f:
        lines = f.read().split("\n")
    problems = []
    n = 0
    for key, value in gold.items():
        if not key.startswith("_") and key != "parameters":
            n += check_entry(key, value, lines, problems)
    return m.group(1), n, problems


def check_all(reference_root):
    report, bad = [], []
    for path in sorted(glob.glob(os.path.join(HERE, "*.json"))):
        log, n, problems = check_fixture(path, reference_root)
        report.append(f"{os.path.basename(path)}: {n} values checked against tests/{log}, {len(problems)} mismatches")
        bad += [f"{os.path.basename(path)}: {p}" for p in problems]
    return report, bad


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    report, bad = check_all(root)
    print("\n".join(report + bad))
    sys.exit(1 if bad else 0)
