#!/bin/bash
# Kernel tuning variants without touching the product library.
#
#   tools/variants.sh build  <name> "<nvcc -D flags>" [EV ...]   # here (no GPU): compiles the instantiation units of the given
#                                                                # state dims (default 4) with the flags, links
#                                                                # tools/micro/_variants/libgpmpc_<name>.so from the product objects
#   tools/variants.sh sass   <name> <kernel-substring>           # here: instruction mix of the variant's hot loops (fp64 vs the rest)
#   tools/variants.sh bench  <name> [bench.py args]              # under gpurun: bench.py against the variant (GPMPC_LIB)
#
# Example (this is how the 3 x 128-thread / 168-register forward kernel was found: a macro around its launch bounds,
# the SASS of the variant inspected here, then ONE short bench on the box):
#   tools/variants.sh build b3 "-DSOME_TUNING_MACRO=3"; tools/variants.sh sass b3 uniform_fwd
#   gpurun -- 'GPMPC_UNI_FWD_THREADS=128 GPMPC_UNI_FWD_CTAS=3 tools/variants.sh bench b3 --steps 2 --no-cpu-baseline'
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
    nvcc $ARCH -shared -o "$OUT/libgpmpc_$name.so" "$CSRC/gpmpc_api.o" "$CSRC/gpmpc_prepare.o" "$CSRC/gpmpc_rollout.o" $objs -lcudart
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
