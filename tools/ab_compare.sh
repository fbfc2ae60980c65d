#!/bin/bash
# Same-box A/B of two builds of libqcnn_b200.so (box-to-box spread of a layer time is ~3-5 %, more than most kernel
# changes are worth): build the alternative here, keep it next to the product library, and let ONE GPU call swap them.
#   (here)  git stash; make -C quantized-cnn_b200 -j8; cp quantized-cnn_b200/libqcnn_b200.so quantized-cnn_b200/alt.so
#           git stash pop; make -C quantized-cnn_b200 -j8
#   gpurun --timeout 400 -- 'bash tools/ab_compare.sh quantized-cnn_b200/alt.so 256'
# Output: gpurun_out/ab_{cur,alt}_{1,2}.log (tools/layer_times.py, two alternating rounds), a one-line summary per run.
set -u
ALT=${1:?path of the alternative libqcnn_b200.so}
B=${2:-256}
LIB=quantized-cnn_b200/libqcnn_b200.so
mkdir -p gpurun_out
cp "$LIB" gpurun_out/.ab_cur.so
for r in 1 2; do
  python tools/layer_times.py --batch "$B" --reps 40 > gpurun_out/ab_cur_$r.log 2>&1
  cp "$ALT" "$LIB"
  python tools/layer_times.py --batch "$B" --reps 40 > gpurun_out/ab_alt_$r.log 2>&1
  cp gpurun_out/.ab_cur.so "$LIB"
done
rm -f gpurun_out/.ab_cur.so
for f in gpurun_out/ab_cur_1.log gpurun_out/ab_alt_1.log gpurun_out/ab_cur_2.log gpurun_out/ab_alt_2.log; do
  echo "$f: $(grep -v autotune "$f" | awk '/^batch/{printf "%s ms/forward |", $3} /^  layer/{printf " %s", $3}')"
done
