#!/bin/bash
# A/B of compile-time variants of the wave kernel (tools/variant_build.py) in one GPU session:
# microseconds per launch of the three ImageNet32 stage shapes (+ the CIFAR batch), and the parity suite of the kernel.
mkdir -p gpurun_out
{
for lib in "" $(ls inverse_flow_b200/lib/exp/libifk_*.so); do
  export IFK_LIBRARY=$lib
  echo "=== variant ${lib:-default}"
  for dims in "100 12 16 16 3 1" "100 24 8 8 3 1" "100 48 4 4 3 1"; do
    timeout 120 python tools/probe_solve.py $dims 2>&1 | grep "^(\|per diagonal\|per launch\|max rel\|transpose\|wait for the image\|write-out\|diagonal loop" | tr '\n' '|' ; echo
  done
done
for lib in $(ls inverse_flow_b200/lib/exp/libifk_*direct*.so inverse_flow_b200/lib/exp/libifk_all3.so); do
  echo "=== wave parity suite with $lib"; IFK_LIBRARY=$lib timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "wave or chain or fused or golden or orientations" 2>&1 | tail -2
done
} > gpurun_out/variant_probe.log 2>&1
tail -70 gpurun_out/variant_probe.log
