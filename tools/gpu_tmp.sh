cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:prune64_lanes_kernel -s 3 -c 1 -o gpurun_out/r2n_lanes python bench.py --steps 2 --warmup 3 --fp64 --no-cpu-baseline --no-host --no-small --no-c5 > gpurun_out/r2n_ncu.log 2>&1; echo ncu_rc=$?
HB2_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --fp64 --no-cpu-baseline --no-host --no-small --no-c5 2> gpurun_out/r2n_debug.err > /dev/null; grep "lanes\|schedule" gpurun_out/r2n_debug.err | sort | uniq -c | head
echo done
