import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, multiprocessing as mp
import config4_lib as c4
from stm32f4_sdr_gps_b200 import Engine, nco_step32
sc=c4.scene(40); sig=c4.signal(sc)
def job(prn):
    from oracle_lib import Reference
    ref=Reference(); chans=ref.channels(1); ref.channel_init(ref.channel_at(chans,0),prn,0)
    return ref.sweep_cells(chans,1,sig[:30],30,-7000,500,29)[0]
if __name__=="__main__":
    eng=Engine(device=0,max_sv=40,ring_ms=64); eng.upload_signal(0,sig)
    for p in range(1,33): eng.set_code_prn(p,p)
    step=np.array([nco_step32(np.float32(4092000-7000+500*b)) for b in range(29)],np.uint32)
    got=eng.sweep(np.arange(1,33),step,0,30)
    with mp.get_context("fork").Pool(16) as pool: want=np.stack(pool.map(job,range(1,33)))
    g=np.stack([got["max"],got["phase"],got["avg"]],axis=-1)
    bad=np.argwhere((g!=want).any(axis=-1))
    print("cells",g.shape,"mismatching",len(bad))
    for b in bad[:10]: print(b, g[tuple(b)], want[tuple(b)])
