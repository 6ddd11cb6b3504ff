#!/usr/bin/env python
"""Regenerates tests/golden/kat.json.

For every known-answer test of the reference (tests.cpp:134-239) this runs the reference's
UNMODIFIED test function -- oracle/_ref/ref_tests_oracle = tests.cpp compiled in place +
oracle/ref_harness_main.cpp + the CPU oracle behind compress()/decompress() -- and records
  * the input the reference's generator produced (dumped by oracle/oracle_dropin.cpp),
  * the compressed words the oracle returned,
  * whether the reference's own ASSERT on its golden array passed.
KAT-1..6 pass bit-exactly.  The two 'wandering' tests compare against a golden that is stale
with respect to the shipped kernel: recorded with ref_assert_passed=false,
used for decode-equivalence and as KAT-7 (the oracle's answer).

Needs /root/reference (run `make -C oracle ref` first); the JSON it writes is committed so
that the GPU box, which has no reference tree, can use it.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_tests_oracle")
TESTS = [
    ("KAT-1", "warpCompressionTest", "tests.cpp:134-152"),
    ("KAT-2", "blockCompressionTest", "tests.cpp:154-164"),
    ("KAT-3", "blockMergeTest", "tests.cpp:166-172"),
    ("KAT-4", "blockMergeWithOnesStartsTest", "tests.cpp:174-185"),
    ("KAT-5", "blockMergeAlternatingTest", "tests.cpp:187-199"),
    ("KAT-6", "blockMergeFinalLiterals", "tests.cpp:201-211"),
    ("KAT-7", "blockMergeWanderingLiterals", "tests.cpp:213-225"),
    ("KAT-8", "multiBlockTest", "tests.cpp:227-239"),
]


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref/ref_tests_oracle first: make -C oracle ref")
    out = []
    for kat, name, cite in TESTS:
        with tempfile.TemporaryDirectory() as d:
            env = dict(os.environ, WAH_DUMP_DIR=d)
            r = subprocess.run([HARNESS, name], env=env, capture_output=True, text=True)
            passed = f"RESULT {name} 1" in r.stdout
            din = np.fromfile(os.path.join(d, "call0_in.u32"), dtype=np.uint32)
            dout = np.fromfile(os.path.join(d, "call0_out.u32"), dtype=np.uint32)
        out.append({
            "kat": kat, "reference_test": name, "cite": cite, "mode": "block1024",
            "ref_assert_passed": passed,
            "input_words": [int(x) for x in din],
            "compressed_words": [int(x) for x in dout],
        })
        print(kat, name, "n =", din.size, "c =", dout.size, "reference ASSERT passed:", passed)
    with open(os.path.join(ROOT, "tests", "golden", "kat.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
