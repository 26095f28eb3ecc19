# phase timing of the path kernels: builds a -DMPO_TAIL_PROF copy of the library ON THE BOX (the shipped .so stays as it is)
set -x
mkdir -p gpurun_out
cd multimodal-path-omic_b200/csrc
cp ../libmpo_b200.so /tmp/libmpo_keep.so
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMPO_TAIL_PROF -c tail_fused.cu -o /tmp/tail_fused_prof.o 2> /tmp/prof_build.log || { tail -20 /tmp/prof_build.log; exit 1; }
objs=$(ls build/*.o | grep -v tail_fused.o)
/usr/local/cuda/bin/nvcc -shared -o ../libmpo_b200.so $objs /tmp/tail_fused_prof.o || exit 1
cd ../..
for m in mcat nacagat; do timeout 200 python scripts/gpu_tail_prof.py $m 32; done 2>&1 | grep -v Warning | tee gpurun_out/tailprof.log
cp /tmp/libmpo_keep.so multimodal-path-omic_b200/libmpo_b200.so
