#!/bin/sh
# Recipe for oracle/_ref/: the reference's own CUDA prototype of the path (feng/ddc/src/ddc_kernel.cu:11-160, launched by
# feng/ddc/src/ddc_host_gpu.py:111-120), compiled for sm_100a FROM WHERE IT LIES under /root/reference.  The source is never
# copied: it is piped through sed into nvcc, and only the cubins land in oracle/_ref/ (git-ignored, shipped to the GPU box).
# The one edit is the sample count, which the prototype hard-codes (`#define N 8192*2`, ddc_kernel.cu:4) and uses to size two
# static __device__ arrays: kernel_ddc_2p14.cubin is the file as shipped, kernel_ddc_2p28.cubin has N = 2^28 (BASELINE
# configs[1]) so that tools/ref_gpu_prototype.py can time the prototype on the headline workload.  The replacement keeps the
# unparenthesised shape of the original (`134217728*2`): the kernel's `cycles/N` relies on it (it expands to cycles/8192*2).
# The prototype is NOT an oracle (FP32 phase, different NCO step: SURVEY 3.3) -- it is a timing baseline only.
set -e
REF=${1:-/root/reference}/feng/ddc/src/ddc_kernel.cu
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$REF" ] || { echo "build_ref.sh: $REF not present (GPU box: the prebuilt cubins are used)"; exit 0; }
mkdir -p "$HERE/_ref"
NVCC=${NVCC:-nvcc}
FLAGS="-diag-suppress 177 -x cu -cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
$NVCC $FLAGS -o "$HERE/_ref/kernel_ddc_2p14.cubin" "$REF"
sed 's/^#define N 8192\*2/#define N 134217728*2/' "$REF" | $NVCC $FLAGS -o "$HERE/_ref/kernel_ddc_2p28.cubin" -
grep -q '^#define N 8192\*2' "$REF" || { echo "build_ref.sh: the N define of the prototype changed"; exit 1; }
echo "oracle/_ref: $(ls "$HERE/_ref")"
