#!/usr/bin/env python
"""CPU campaigns (no GPU), the long form of two tests:

  oracles  the streaming restatement (oracle/htslib_pileup.py under the restated vafator column functions) against the
           direct per-pair C restatement on seeded noisy synthetic reads (tests/test_oracle_pins.py)
  fetch    the windows of vafator_b200/fetch.py against the span the reference fetches, through the C oracle, on
           low-coverage reads full of long reference skips (tests/test_fetch.py)

    python tests/fuzz_oracles.py [--seconds 120] [--what oracles,fetch]
"""
import argparse
import pathlib
import shutil
import sys
import tempfile
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import test_fetch  # noqa: E402
import test_oracle_pins  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--what", default="oracles,fetch")
    ap.add_argument("--first-seed", type=int, default=1000)
    a = ap.parse_args()
    bad = 0
    for what in a.what.split(","):
        t, seed, ok = time.time(), a.first_seed, 0
        while time.time() - t < a.seconds / len(a.what.split(",")):
            try:
                if what == "oracles":
                    test_oracle_pins.test_two_formulations_agree_on_random_reads(seed)
                else:
                    d = pathlib.Path(tempfile.mkdtemp(prefix="vfk_fetch"))
                    try:
                        test_fetch.test_windows_on_low_coverage_reads_with_long_skips(d, seed)
                    finally:
                        shutil.rmtree(d, ignore_errors=True)
                ok += 1
            except AssertionError as e:
                if what == "fetch" and "extended" in str(e):
                    ok += 1  # no window of this seed needed an extension: nothing wrong with the fetch
                else:
                    bad += 1
                    print("%s: seed %d: %s" % (what, seed, str(e)[:400]), flush=True)
            seed += 1
        print("%s: %d seeds agree (%d..%d), %.0f s" % (what, ok, a.first_seed, seed - 1, time.time() - t), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
