#!/bin/bash
# Kernel tuning variants without touching the product library.
#
#   tools/variants.sh build  <name> "<nvcc -D flags>" [EV ...]   # here (no GPU): compiles the instantiation units of the given
#                                                                # state dims (default 4) with the flags, links
#                                                                # tools/micro/_variants/libgpmpc_<name>.so from the product objects
#   tools/variants.sh build-all <name> "<nvcc -D flags>"         # here: ALL units with the flags (macros that change the
#                                                                # shared-memory layout must reach the host code too)
#   tools/variants.sh sass   <name> <kernel-substring>           # here: instruction mix of the variant's hot loops (fp64 vs the rest)
#   tools/variants.sh bench  <name> [bench.py args]              # under gpurun: bench.py against the variant (GPMPC_LIB)
#
# Example (this is how the 3 x 128-thread / 168-register forward kernel was found: a macro around its launch bounds,
# the SASS of the variant inspected here, then ONE short bench on the box):
#   tools/variants.sh build b3 "-DSOME_TUNING_MACRO=3"; tools/variants.sh sass b3 uniform_fwd
#   gpurun -- 'GPMPC_UNI_FWD_THREADS=128 GPMPC_UNI_FWD_CTAS=3 tools/variants.sh bench b3 --steps 2 --no-cpu-baseline'
# Prepared for the next round -- the reverse-sweep kernel as three 128-thread CTAs per SM with 168 registers:
#   tools/variants.sh build-all bw3 "-DUNI_BWD_MAXT_ALL=128 -DUNI_BWD_MINCTAS_ALL=3 -DGPMPC_BWD_NO_CST"
#   gpurun -- 'GPMPC_UNI_PREMAT=0 GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=3 tools/variants.sh bench bw3 --steps 2 --no-cpu-baseline'
# (big batches only: the cluster launches of small batches use 256 threads, beyond the variant's launch bounds)
# The variant libraries are git-ignored (*.so) but travel to the GPU box; each adds ~37 MB to the push, so delete
# tools/micro/_variants/ when done.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC="$ROOT/data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200/csrc"
OUT="$ROOT/tools/micro/_variants"
ARCH="-gencode arch=compute_100a,code=sm_100a"
cmd=$1; name=$2
case "$cmd" in
  build)
    flags=$3; shift 3 || true
    evs=${@:-4}
    mkdir -p "$OUT"
    make -C "$CSRC" -j 16 > /dev/null
    objs=""
    for n in 1 2 3 4 5 6 7 8; do
      if [[ " $evs " == *" $n "* ]]; then
        nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v $flags -c "$CSRC/gpmpc_inst_ev$n.cu" \
             -o "$OUT/${name}_ev$n.o" 2> "$OUT/${name}_ev$n.ptxas.log"
        grep -A2 "uniform_\|rollout_kernel" "$OUT/${name}_ev$n.ptxas.log" | grep "Compiling\|registers" | sed 's/ptxas info    : //' | cut -c1-150
        objs="$objs $OUT/${name}_ev$n.o"
      else
        objs="$objs $CSRC/gpmpc_inst_ev$n.o"
      fi
    done
    nvcc $ARCH -shared -o "$OUT/libgpmpc_$name.so" "$CSRC/gpmpc_api.o" "$CSRC/gpmpc_prepare.o" "$CSRC/gpmpc_rollout.o" "$CSRC/gpmpc_optim.o" $objs -lcudart
    ls -la "$OUT/libgpmpc_$name.so"
    ;;
  build-all)
    flags=$3
    mkdir -p "$OUT"
    objs=""
    for src in gpmpc_api gpmpc_prepare gpmpc_rollout gpmpc_optim gpmpc_inst_ev1 gpmpc_inst_ev2 gpmpc_inst_ev3 gpmpc_inst_ev4 gpmpc_inst_ev5 \
               gpmpc_inst_ev6 gpmpc_inst_ev7 gpmpc_inst_ev8; do
      ( nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v $flags -c "$CSRC/$src.cu" -o "$OUT/${name}_${src#gpmpc_inst_}.o" \
             2> "$OUT/${name}_${src#gpmpc_inst_}.ptxas.log" || { cat "$OUT/${name}_${src#gpmpc_inst_}.ptxas.log" | grep -i error; exit 1; } ) &
      objs="$objs $OUT/${name}_${src#gpmpc_inst_}.o"
    done
    wait
    nvcc $ARCH -shared -o "$OUT/libgpmpc_$name.so" $objs -lcudart
    grep -A2 "uniform_bwd_kernelILi4" "$OUT/${name}_ev4.ptxas.log" | grep "registers\|spill" | sed 's/ptxas info    : //'
    ls -la "$OUT/libgpmpc_$name.so"
    ;;
  sass)
    for o in "$OUT/${name}"_ev*.o; do
      for fn in $(cuobjdump -sass "$o" | grep "Function :" | grep "$3" | awk '{print $3}'); do
        echo "## $fn ($(basename $o))"
        cuobjdump -sass -fun "$fn" "$o" | python3 "$ROOT/tools/sass_loops.py"
      done
    done
    ;;
  bench)
    shift 2
    GPMPC_LIB="$OUT/libgpmpc_$name.so" python "$ROOT/bench.py" "$@"
    ;;
  *) sed -n 2,16p "$0"; exit 1;;
esac
