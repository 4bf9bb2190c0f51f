#!/bin/bash
# Produces the records profiles/README.md quotes, on a 1-GPU box (run under gpurun; outputs in gpurun_out/):
#   bash tools/gpu_records.sh [tests|bench|ref|sweep|ncu|launches|sanitize|scale]...     (default: tests bench ref)
# Multi-GPU records: torchrun ... bench.py --gpus N (see README.md) and tests/multigpu_check.py.
set -x
mkdir -p gpurun_out
what=${@:-tests bench ref}
for w in $what; do case $w in
tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
       python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ;;
bench) timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json ;;
ref)   timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_n1.json 2> gpurun_out/ref_n1.err; cut -c1-400 gpurun_out/ref_n1.json ;;
sweep) timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-strong --no-cpu --sweep > gpurun_out/sweep.json 2> gpurun_out/sweep.err ;;
ncu)   NCU="ncu --set full --import-source on --clock-control none"
       timeout 400 $NCU -k regex:k_clip_fast -s 1 -c 1 -o gpurun_out/clip -f python tools/prof_one.py 4096 1 > gpurun_out/ncu.log 2>&1
       timeout 400 $NCU -k regex:k_matvec -s 4 -c 1 -o gpurun_out/mv -f python tools/prof_one.py 4096 2 >> gpurun_out/ncu.log 2>&1
       timeout 400 $NCU -k regex:k_cg_update -s 4 -c 2 -o gpurun_out/upd -f python tools/prof_one.py 4096 2 >> gpurun_out/ncu.log 2>&1
       timeout 400 $NCU -k regex:'k_pipe_copy|k_pipe_deg|k_pipe_hdr' -c 3 -o gpurun_out/pipe -f python tools/e2e_trace.py 4096 >> gpurun_out/ncu.log 2>&1
       python tools/ncu_summary.py gpurun_out/clip.ncu-rep gpurun_out/mv.ncu-rep gpurun_out/upd.ncu-rep gpurun_out/pipe.ncu-rep > gpurun_out/ncu_summary.txt ;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
       python bench.py --steps 1 --warmup 1 --no-e2e --no-strong --no-cpu --no-shuffle --side 2048 > gpurun_out/launches.out 2>&1 ;;
sanitize) bash tools/sanitize.sh gpurun_out ;;
scale) timeout 1500 python tools/configs_at_scale.py > gpurun_out/configs_at_scale.json 2> gpurun_out/configs.err ;;
stats) LV_CLIP_STATS=1 python tools/prof_one.py 4096 1 2>&1 | grep "clip stats"; python tools/e2e_trace.py 4096 2>&1 | grep "^step\|job" | tail -4 ;;
esac; done
