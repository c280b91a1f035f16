#!/usr/bin/env python
"""Stop a job started with poseidon_b200.tools.launch (reference: scripts/kill_caffe.py, which runs `killall caffe_main` on
every host of the hostfile).  Signals exactly the recorded processes:  scripts/kill_job.py output/run [--force]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseidon_b200.tools import launch  # noqa: E402

if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    sys.exit(launch.main(["kill", "--run_dir", sys.argv[1]] + sys.argv[2:]))
