import sys, time
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np
import config4_lib as c4
from stm32f4_sdr_gps_b200 import Engine, Channels, Receiver
n_trk=10000
sc=c4.scene(n_trk+700); sig=c4.signal(sc)
eng=Engine(device=0,max_sv=40,ring_ms=sc.n_ms); eng.upload_signal(0,sig)
for log in (True, False):
    ch=Channels(c4.SEARCHED); rx=Receiver(eng,ch); rx.set_slot_walk(True)
    rep=rx.cold_start(0,sweeps=3); t=rep["ms_next"]
    rx.track_run(t,160,log=log)
    t0=time.perf_counter(); rx.track_run(t+160,n_trk-160,log=log); t1=time.perf_counter()
    print("32 channels log",log,"%.2f ms"%((t1-t0)*1e3), rx.loop_stats())
    found=[i for i in range(32) if ch.snapshot(i).acq_state==9]
    rx.close(); ch.free()
