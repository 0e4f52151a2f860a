import sys
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np
import config4_lib as c4
from oracle_lib import Reference
from stm32f4_sdr_gps_b200 import Engine
n_trk=int(sys.argv[1]) if len(sys.argv)>1 else 1500
sc=c4.scene(n_trk+700); sig=c4.signal(sc)
print("present",[ (s.prn, round(s.doppler_hz), round(s.code_phase_samples/8)) for s in sc.sats])
eng=Engine(device=0,max_sv=40,ring_ms=sc.n_ms); eng.upload_signal(0,sig)
tm={}
ch,rx,rep,logs=c4.product(eng,sig,c4.SEARCHED,n_trk,timers=tm)
print(rep, tm)
refs=c4.reference_all(sig,c4.SEARCHED,rep,n_trk)
plain=lambda v: bytes(v) if hasattr(v,'__len__') else v
for i,prn in enumerate(c4.SEARCHED):
    a=ch.snapshot(i); b=type(a).from_buffer_copy(refs[i][0])
    d=[(n,plain(getattr(a,n)),plain(getattr(b,n))) for n,_ in a._fields_ if plain(getattr(a,n))!=plain(getattr(b,n))]
    if a.acq_state!=1 or d:
        iqd=np.argwhere((logs[0][:,i,:]!=refs[i][1]).any(axis=1))
        print(prn,"acq",a.acq_state,b.acq_state,"trk",a.trk_state,"f",a.found_freq_offset_hz,"cp",a.found_code_phase,"diff",d[:6],"first iq diff row",iqd[:1].tolist())
import ctypes as C
for i,prn in enumerate(c4.SEARCHED):
    a=ch.snapshot(i)
    if a.acq_state==9:
        s=rx.sync_status(i)
        print(prn,"sync",s.bit_period_found,"refined",s.bit_edge_refined,"phase",s.slot_phase,"walks",s.walks,"pending",s.walk_pending,"words",s.words_ok,"subframes",s.subframes,"right_period_cnt",a.right_period_cnt,"cpf",np.uint32(a.code_phase_fine_bits).view(np.float32), "idle rows", int((~logs[0][200:,i,:].any(axis=1)).sum()))
