#!/bin/bash
# NVLink byte counters and durations of the DYNAMIC x sweep on a real 2-GPU ring: rank 0 runs under ncu (a handful of
# metrics, few replay passes), rank 1 runs plain.   usage (2-GPU box): bash tools/ring_ncu.sh
# (under ncu's kernel replay the peer may hit its 2 s halo watchdog: irrelevant here, only rank 0's counters are read)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ring_wrap.sh <<'EOS'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,launch__registers_per_thread,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
       --clock-control none -k regex:sweep_ -c 16 --csv --log-file gpurun_out/ring_ncu_rank0.csv python tools/ring_steps.py 4
else
  exec python tools/ring_steps.py 4
fi
EOS
chmod +x /tmp/ring_wrap.sh
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 --no-python /tmp/ring_wrap.sh > gpurun_out/ring_ncu.log 2>&1
tail -5 gpurun_out/ring_ncu.log
