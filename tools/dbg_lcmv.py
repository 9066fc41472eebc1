import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import beamform_b200 as bf
from beamform_b200.synth import synth_batch
from oracle_lib import Oracle
from test_gpu_parity import run_device, oracle_with_flags, H
for algo, ints, events in [("lcmv",(80.0,-60.0,150.0),[]), ("lcmv",(80.0,),[]), ("lcmv",(),[]),
      ("lcmv",(80.0,-60.0,150.0),[(20, "theta", 20.0), (40, "interf", 2, -55.0), (60, "interf", 4, 120.0), (80, "interf", 1, 119.5), (90, "interf", 0, 10.0)])]:
    cfg = bf.make_config(algo, mics="circ8", initial_angle=0.0, interferers=ints)
    x = synth_batch(bf.GEOMETRIES["circ8"], 2, 100 * H, seed=31)
    ref, sel, _ = oracle_with_flags(cfg, x, events=events)
    got, flags, b = run_device(cfg, x, events=events)
    print(algo, ints, "events" if events else "", "total", np.linalg.norm(got-ref)/np.linalg.norm(ref))
    for blk in range(0,100,10):
        sl = slice(blk*H,(blk+10)*H)
        r = ref[:,sl]; g=got[:,sl]
        print("  hops %3d-%3d  rel %.3e  refnorm %.3e"%(blk,blk+10,np.linalg.norm(g-r)/max(np.linalg.norm(r),1e-30), np.linalg.norm(r)))
