# usage: scripts/build_variant.sh NAME   -- rebuilds the chi=64 object from the working tree and links build/variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMDOPT_CHI=64 -c -o build/variants/inst_chi64_$1.o mdopt_b200/csrc/inst.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib$1.so build/inst_chi8.o build/inst_chi16.o build/inst_chi32.o build/variants/inst_chi64_$1.o build/inst_chi128.o build/inst_chi256.o build/capi.o
ls -la build/variants/lib$1.so
