"""Per-pass wall times of the GPU-decode file loop in a fresh, torch-free process (768 synthetic 640x480 frame pairs written with
cv2, DECODE_AHEAD 3 / 2 / 3, four passes each).  Round-2 record on a B200 box: [400, 158, 158, 158] / [220, 167, 167, 224] /
[163, 160, 153, 157] ms -- the loop is stable after its first (allocating) pass; see DESIGN.md section 6."""
import os, sys, time, json, tempfile, shutil
t00=time.perf_counter()
sys.path.insert(0, os.getcwd())
import numpy as np, cv2
import otslam_b200.o3d_compat as o3d
from otslam_b200 import pipeline, capture
T_FIX = np.array([[0,0,1,0],[-1,0,0,0],[0,-1,0,0],[0,0,0,1.0]])
H,W,n=480,640,768
rng=np.random.default_rng(0)
yy,xx=np.mgrid[0:H,0:W]
base=tempfile.mkdtemp(prefix='diag_')
for k in range(n):
    d=(900+200*np.sin(xx/80.0+k*0.01)+100*np.cos(yy/60.0)).astype(np.uint16)
    c=np.dstack([(xx+k)%256,(yy*2)%256,((xx+yy)//2)%256]).astype(np.uint8)
    pose=np.eye(4); pose[0,3]=0.001*k
    capture.save_frame(base,'Object_0',k+1,c,d,pose)
tr=[(f'{base}/color/Object_0_{j}.jpg',f'{base}/depth/Object_0_{j}.png',f'{base}/poses/Object_0_{j}.txt',j) for j in range(1,n+1)]
intr=o3d.camera.PinholeCameraIntrinsic(W,H,565.6,565.6,320.5,240.5)
vol=o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=0.01,sdf_trunc=0.04,color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)
out={'setup_s':time.perf_counter()-t00}
for ahead in (3,2,3):
    pipeline.DECODE_AHEAD=ahead
    ts=[]
    for p in range(4):
        vol.reset(); t0=time.perf_counter(); pipeline.integrate_files(vol,tr,intr,T_FIX); ts.append(round(1e3*(time.perf_counter()-t0),1))
    out[f'ahead{ahead}_{len(out)}']=ts
    print(json.dumps(out),flush=True)
shutil.rmtree(base,ignore_errors=True)
