"""Merges the per-process parity logs of a GPU test run (tests/conftest.py: parity_worst.<pid>.json — the worst pose
difference engine vs oracle each test's assertions saw) into one table.

    python tools/summarize_parity.py gpurun_out profiles/r02_parity_worst.json
"""
import glob
import json
import os
import sys


def main(src, dst):
    merged = {}
    for f in sorted(glob.glob(os.path.join(src, "parity_worst.*.json"))):
        for test, v in json.load(open(f)).items():
            m = merged.setdefault(test, {"max_translation_m": 0.0, "max_rotation_rad": 0.0, "frames_compared": 0})
            m["max_translation_m"] = max(m["max_translation_m"], v["max_translation_m"])
            m["max_rotation_rad"] = max(m["max_rotation_rad"], v["max_rotation_rad"])
            m["frames_compared"] += v["frames_compared"]
    json.dump(dict(sorted(merged.items())), open(dst, "w"), indent=1)
    for test, v in sorted(merged.items()):
        print("%-95s %9.2e m %9.2e rad  (%d frames)" % (test[-95:], v["max_translation_m"], v["max_rotation_rad"], v["frames_compared"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
