# one ncu capture of the kernel matching $1 (regex) for bench config $2 with --nb $3 (after a plain run of the same command)
K=$1; CFG=$2; NB=$3; OUT=$4; shift 4
CMD="python bench.py --config $CFG --nb $NB --steps 1 --warmup 3 --no-cpu-baseline --no-extras $@"
timeout 200 $CMD > gpurun_out/ncu_plain_$OUT.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$OUT.log
