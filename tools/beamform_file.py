#!/usr/bin/env python
"""Offline file driver: the stand-in for the JACK/ROS transport (rosjack.cpp) when beamforming recorded audio.

    python tools/beamform_file.py --algo mvdr --config beamform_config.yaml --in mics.wav --out beam.wav [--theta 20] [--hop 512]
                                  [--set key=value ...] [--theta-at HOP:DEG ...] [--interf-at HOP:ID:DEG ...]

`--config` is the reference's beamform_config.yaml, unchanged (microphone geometry, initial angle, interferers); `--set` takes the
launch-file <rosparam> keys of the node (launch/<algo>.launch values are the defaults).  The WAV holds one channel per microphone
(PCM16/PCM32/float32); the output is mono float32 at the same rate.  The JACK period (`--hop`) fixes the frame size (2 x hop).
Everything runs through the C ABI of libbeamform_b200.so (bf_config_load_yaml / bf_create / bf_process_batch)."""
import argparse
import os
import sys

import numpy as np
from scipy.io import wavfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import beamform_b200 as bf  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--algo", required=True, choices=sorted(bf.ALGOS))
    ap.add_argument("--config", required=True, help="beamform_config.yaml")
    ap.add_argument("--in", dest="inp", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--theta", type=float, default=None, help="overrides initial_angle of the yaml")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE")
    ap.add_argument("--theta-at", action="append", default=[], metavar="HOP:DEG")
    ap.add_argument("--interf-at", action="append", default=[], metavar="HOP:ID:DEG")
    a = ap.parse_args(argv)
    sr, wav = wavfile.read(a.inp)
    if wav.ndim == 1:
        wav = wav[:, None]
    if wav.dtype == np.int16:
        x = wav.astype(np.float32) / 32768.0
    elif wav.dtype == np.int32:
        x = wav.astype(np.float32) / 2147483648.0
    else:
        x = wav.astype(np.float32)
    params = {}
    for kv in a.set:
        k, v = kv.split("=", 1)
        params[k] = (v.lower() == "true") if v.lower() in ("true", "false") else float(v)
    cfg = bf.load_yaml_config(a.algo, a.config, hop=a.hop, sample_rate=sr, **params)
    if a.theta is not None:
        cfg.initial_angle = a.theta
    if x.shape[1] < cfg.n_mics:
        sys.exit("the file has %d channels, the configuration %d microphones" % (x.shape[1], cfg.n_mics))
    n_hops = x.shape[0] // a.hop          # JACK delivers whole periods only
    xin = np.ascontiguousarray(x[:n_hops * a.hop, :cfg.n_mics].T)[None]   # [1][M][L]
    events = [(int(h), "theta", float(d)) for h, d in (s.split(":") for s in a.theta_at)]
    events += [(int(h), "interf", int(i), float(d)) for h, i, d in (s.split(":") for s in a.interf_at)]
    y = bf.Beamformer(cfg, n_streams=1).process(xin, events=events)[0]
    wavfile.write(a.out, sr, y.astype(np.float32))
    clipped = int((np.abs(y) >= 1.0).sum())   # output_to_rosjack warns and does not clip (rosjack.cpp:372-374)
    print("%s: %d microphones, %d hops of %d at %d Hz -> %s%s" % (a.algo, cfg.n_mics, n_hops, a.hop, sr, a.out,
                                                                 " (%d samples at or above full scale)" % clipped if clipped else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
