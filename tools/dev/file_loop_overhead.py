"""Host-side cost of the GPU-decode file loop (pipeline._integrate_files_gpu) WITHOUT a GPU: stand-in decoders that sleep 45 ms per
256-frame chunk (the measured device time of a chunk) and 6 ms per integrate, real pose files.  Prints the event timeline of three
passes over 768 frames; ideal = 45 + 3 x 6 = 63 ms.  (Before the pose files were parsed by the library: ~200 ms, 16 interpreter
threads fighting for the GIL; after: 68 ms.)"""
import os, sys, time, threading, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import otslam_b200.o3d_compat as o3d
from otslam_b200 import pipeline, capture
T_FIX=np.array([[0,0,1,0],[-1,0,0,0],[0,-1,0,0],[0,0,0,1.0]])
n=768
base=tempfile.mkdtemp(); os.makedirs(base+'/poses')
for k in range(n):
    open(f'{base}/poses/Object_0_{k+1}.txt','w').write(capture.pose_text(np.eye(4)+0.001*k*np.eye(4,k=3)))
tr=[(f'{base}/color/Object_0_{j}.jpg',f'{base}/depth/Object_0_{j}.png',f'{base}/poses/Object_0_{j}.txt',j) for j in range(1,n+1)]
T0=[0]; ev=[]
def log(m): ev.append((round(1e3*(time.perf_counter()-T0[0]),1), m))
class Fake:
    def decode_files(self,c,d):
        log('dec start'); time.sleep(0.045); log('dec end'); return np.zeros(len(c),np.int32), np.zeros(len(c),np.int32)
    def profile(self): return {"compressed_bytes":0}
    def integrate(self,*a): log('int start'); time.sleep(0.006); log('int end')
    def close(self): pass
def acquire(count,h,w,f,dev):
    pipeline._decoder_lock = pipeline._decoder_lock or threading.Lock()
    return [Fake() for _ in range(count)], ('fake',count)
pipeline._acquire_decoders=acquire
class Vol:
    class _vol: device=0
intr=o3d.camera.PinholeCameraIntrinsic(640,480,565.6,565.6,320.5,240.5)
pipeline.DECODE_AHEAD=3
for p in range(3):
    ev.clear(); T0[0]=time.perf_counter(); pipeline._integrate_files_gpu(Vol,tr,intr,T_FIX,1000.0,3.0,False,None,None); log('done')
print(ev)
shutil.rmtree(base)
