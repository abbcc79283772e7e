# usage: [CHI=64] scripts/build_variant.sh NAME [extra nvcc flags]  -- rebuilds the dense-engine object of one capacity
# from the working tree with the extra flags and links build/variants/libNAME.so from it and the other objects of the
# regular build (select it at run time with MDOPT_B200_LIB=build/variants/libNAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
chi=${CHI:-64}
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMDOPT_CHI=$chi "$@" -c -o build/variants/inst_chi${chi}_$name.o mdopt_b200/csrc/inst.cu
objs=$(ls build/inst_chi*.o build/mono_chi*.o build/capi.o | grep -v "inst_chi$chi.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib$name.so $objs build/variants/inst_chi${chi}_$name.o
ls -la build/variants/lib$name.so
