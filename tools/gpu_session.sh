#!/bin/bash
# One GPU lease = one call of this script under gpurun; every step runs under its own `timeout`, writes into gpurun_out/
# and never stops the ones after it.   usage: tools/gpu_session.sh <tag> <step> [<step> ...]
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
export SWB200_FEED_TIMEOUT_MS=${SWB200_FEED_TIMEOUT_MS:-8000}
{ nvidia-smi -L; nproc; free -g | head -2; numactl -H 2>/dev/null | head -3; nvidia-smi topo -m 2>/dev/null | head -12; } > "$out/box.txt" 2>&1
for step in "$@"; do
  t0=$(date +%s)
  case $step in
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1 ;;
    parity)   timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu > "$out/parity.log" 2>&1 ;;
    gputests) timeout 2400 python -m pytest tests -x -q -m gpu > "$out/gputests.log" 2>&1 ;;
    bench)    timeout 900 python bench.py > "$out/bench.json" 2> "$out/bench.err" ;;
    benchq)   timeout 300 python bench.py --quick --no-cpu-baseline --steps 10 > "$out/benchq.json" 2> "$out/benchq.err" ;;
    benchq_plain) SWB200_PACK_STREAM=0 timeout 300 python bench.py --quick --no-cpu-baseline --steps 10 > "$out/benchq_plainstores.json" 2> "$out/benchq_plainstores.err" ;;
    benchq_tl) SWB200_FEED_TIMELINE=1 timeout 300 python bench.py --quick --no-cpu-baseline --steps 6 > "$out/benchq_timeline.json" 2> "$out/benchq_timeline.err" ;;
    benchq_big) SWB200_FEED_FIRST_TILES=8 SWB200_FEED_TIMELINE=1 timeout 300 python bench.py --quick --no-cpu-baseline --steps 6 > "$out/benchq_first8.json" 2> "$out/benchq_first8.err" ;;
    lanes_matrix)
              for cfg in "2 15" "4 15" "8 15" "2 7" "4 7" "2 11"; do set -- $cfg
                echo "== lane tiles $1, pack threads $2" >> "$out/lanes_matrix.txt"
                SWB200_FEED_LANE_TILES=$1 SWB200_FEED_TIMELINE=1 timeout 200 python bench.py --quick --no-cpu-baseline --no-plain-e2e --steps 6 --pack-threads $2 2>&1 >/dev/null | grep timeline | sed -n '6,8p' >> "$out/lanes_matrix.txt"
              done
              echo "== plain stores" >> "$out/lanes_matrix.txt"
              SWB200_PACK_STREAM=0 SWB200_FEED_TIMELINE=1 timeout 200 python bench.py --quick --no-cpu-baseline --no-plain-e2e --steps 6 2>&1 >/dev/null | grep timeline | sed -n '6,8p' >> "$out/lanes_matrix.txt" ;;
    zcprobe)  timeout 200 python tools/zero_copy_probe.py > "$out/zero_copy_probe.json" 2> "$out/zero_copy_probe.err" ;;
    refarm)   timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$out/reference_arm.json" 2> "$out/reference_arm.err" ;;
    sgtests)  timeout 900 python -m pytest tests/test_semiglobal_gpu.py -x -q -m gpu > "$out/sgtests.log" 2>&1 ;;
    sgbench)  timeout 600 python bench.py --workload semiglobal --no-cpu-baseline --steps 5 > "$out/sgbench.json" 2> "$out/sgbench.err" ;;
    sgbench_tl) SWB200_SG_TIMELINE=1 timeout 600 python bench.py --workload semiglobal --no-cpu-baseline --steps 6 > "$out/sgbench_tl.json" 2> "$out/sgbench_tl.err" ;;
    sgbench_depths) for dd in ${SG_DEPTHS:-8 4}; do echo "== forward depth $dd" >> "$out/sgbench_depths.txt"; SWB200_SG_FWD_DEPTH=$dd SWB200_SG_TIMELINE=1 timeout 300 python bench.py --workload semiglobal --no-cpu-baseline --steps 2 2>&1 >/dev/null | grep "sg pipeline" | tail -17 >> "$out/sgbench_depths.txt"; done ;;
    sgbench_large) SWB200_SG_TIMELINE=1 timeout 900 python bench.py --workload semiglobal --no-cpu-baseline --steps 3 --pairs 151552 > "$out/sgbench_large.json" 2> "$out/sgbench_large.err" ;;
    sgbench_plain) SWB200_SG_PIPE=0 timeout 600 python bench.py --workload semiglobal --no-cpu-baseline --steps 5 > "$out/sgbench_plain.json" 2> "$out/sgbench_plain.err" ;;
    kbench)   timeout 120 tools/kbench > "$out/kbench.jsonl" 2>&1 ;;
    bench2|bench4|bench8)
              n=${step#bench}
              timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n > "$out/bench_n$n.json" 2> "$out/bench_n$n.err" ;;
    ncu_launches) SWB200_FEED_NO_RELAY=1 SWB200_FEED_TIMEOUT_MS=60000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sw_ -c 400 --csv --log-file "$out/ncu_launches.csv" python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 > "$out/ncu_launches.out" 2>&1 ;;
    ncu_pair) timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_pair_kernel -s 200 -c 1 -o "$out/ncu_full_sw_pair_kernel" tools/speedtest_b200 3000 > "$out/ncu_pair.out" 2>&1 ;;
    packbench) timeout 300 tools/packbench 64 > "$out/packbench.jsonl" 2>&1 ;;
    speedtest) timeout 120 tools/speedtest_b200 20000 > "$out/speedtest.json" 2>&1 ;;
    speedtest_gaps) for g in 0 300 500 700 1000 1500; do SWB200_PAIR_POLL_GAP_NS=$g timeout 120 tools/speedtest_b200 20000 2>&1 | sed "s/^{/{\"poll_gap_ns\": $g, /" >> "$out/speedtest_gaps.jsonl"; done ;;
    ncu_full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_kernel -s 3 -c 1 -o "$out/ncu_full_sw_kernel" python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 > "$out/ncu_full.out" 2>&1
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_feed_kernel -s 2 -c 1 -o "$out/ncu_full_sw_feed_kernel" python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 >> "$out/ncu_full.out" 2>&1 ;;
    *) echo "unknown step $step" ;;
  esac
  echo "$step rc=$? $(( $(date +%s) - t0 )) s" | tee -a "$out/steps.txt"
done
tail -n 3 "$out"/*.log 2>/dev/null | tail -n 40
