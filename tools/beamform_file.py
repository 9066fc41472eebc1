#!/usr/bin/env python
"""Offline file driver: thin wrapper around the C++ tool tools/bf_offline (built by `python -m beamform_b200.build`), the
stand-in for the JACK/ROS transport (rosjack.cpp) when beamforming recorded audio.

    python tools/beamform_file.py --algo mvdr --config beamform_config.yaml [--launch mvdr.launch] --in mics.wav --out beam.wav
                                  [--theta 20] [--hop 512] [--set key=value ...] [--theta-at HOP:DEG ...] [--interf-at HOP:ID:DEG ...]

Without --launch the launch-file values of the node (beamform_b200/tables.py: LAUNCH_PARAMS, taken from launch/<algo>.launch) are
passed as --set defaults, so the tool behaves like `roslaunch beamform <algo>.launch`.  All arguments go to bf_offline unchanged."""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    from beamform_b200 import build, tables
    exe = build.build_tools()
    extra = []
    if "--launch" not in argv and "--algo" in argv:
        algo = argv[argv.index("--algo") + 1]
        for k, v in tables.LAUNCH_PARAMS.get(algo, {}).items():
            extra += ["--set", "%s=%s" % (k, ("true" if v else "false") if isinstance(v, bool) else repr(float(v)))]
    # launch defaults first: later --set arguments of the caller override them
    i = argv.index("--algo") + 2 if "--algo" in argv else 0
    return subprocess.run([exe] + argv[:i] + extra + argv[i:]).returncode


if __name__ == "__main__":
    sys.exit(main())
