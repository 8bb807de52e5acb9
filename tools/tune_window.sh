# ad-hoc sweep of window-kernel variants (IFK_WINDOW_CFG=cc,nv,cluster,rp) on a few large shapes
run() { IFK_WINDOW_CFG=$1 timeout 60 python tools/microbench.py --solve-only --iters 3 --shape $2 2>&1 | grep inverse | cut -c1-175; }
for cfg in 4,3,1,1 4,3,1,2 6,3,1,1 6,3,1,2 8,3,1,2 3,6,1,2 4,6,1,2; do run $cfg 512,48,32,32,3,1; done
for cfg in 6,6,2,1 8,3,2,2 6,3,2,2 4,6,2,2 4,3,2,2; do run $cfg 64,48,32,32,3,1; done
for cfg in 4,9,2,1 2,9,2,2 4,5,4,2 4,5,2,2; do run $cfg 64,48,32,32,5,1; done
for cfg in 3,6,4,1 3,6,4,2 4,6,4,2 4,3,4,2 6,3,4,2; do run $cfg 64,96,32,32,3,1; done
for cfg in 12,3,1,1 8,3,1,2 6,3,1,2; do run $cfg 512,12,64,64,3,1; done
