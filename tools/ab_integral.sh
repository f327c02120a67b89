#!/bin/bash
# gs_integral strips kernel variants (streaming stores, register caps) on the ops (64 x 4096^2) and c4 (256 x UHD) shapes
run() { for wl in ops c4; do extra=""; [ $wl = c4 ] && extra="--steps 2 --warmup 1"; [ $wl = ops ] && extra="--steps 10 --warmup 3"; timeout 300 python bench.py --workload $wl $extra --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k='gs_integral'; print('%-12s %-4s %.3f ms  %.3f' % ('$1', '$wl', d['kernels'][k]['ms'], d['kernels'][k]['frac']))"; done; }
run default
for v in iscs isminb5 isminb6cs; do [ -f grayskull_b200/libv_$v.so ] && GS_B200_LIB=$PWD/grayskull_b200/libv_$v.so run $v; done
GS_B200_INTEGRAL=bands run bands
