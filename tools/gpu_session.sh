#!/bin/bash
# One gpurun call = several measurements; everything lands in gpurun_out/ (merged back by gpurun).
# usage: tools/gpu_session.sh <tag> [steps...]   steps: pytest bench ncu_list ncu_full sanitizer probe
tag=$1; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_smi.txt 2>&1
for step in "$@"; do
  case $step in
    pytest)
      # one process per file, each under its own timeout: a hang costs one file, not the session; slowest tests are listed
      for f in ${PYTEST_FILES:-tests/test_gpu_misc.py tests/test_gpu_compress.py tests/test_gpu_inflate.py tests/test_gpu_foreign.py tests/test_gpu_framing.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py}; do
        echo "=== $f $(date +%T)" >> $out/${tag}_pytest.log
        timeout ${PYTEST_TIMEOUT:-600} python -m pytest $f -m gpu -k "${PYTEST_K:-test}" --maxfail=8 -v --durations=6 -o faulthandler_timeout=150 >> $out/${tag}_pytest.log 2>&1; echo "exit $? $(date +%T)" >> $out/${tag}_pytest.log
      done ;;
    bench)
      timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?" >> $out/${tag}_bench.err ;;
    bench_args)
      timeout 900 python bench.py ${BENCH_ARGS} > $out/${tag}_bencha.json 2> $out/${tag}_bencha.err; echo "bench exit $?" >> $out/${tag}_bencha.err ;;
    bench_multi)
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} ${BENCH_ARGS} > $out/${tag}_bench_${NGPU:-2}gpu.json 2> $out/${tag}_bench_${NGPU:-2}gpu.err; echo "bench exit $?" >> $out/${tag}_bench_${NGPU:-2}gpu.err ;;
    bench_quick)
      timeout 600 python bench.py --steps 3 --warmup 3 --skip config5,dropin > $out/${tag}_benchq.json 2> $out/${tag}_benchq.err; echo "bench exit $?" >> $out/${tag}_benchq.err ;;
    ncu_list)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
        python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --skip config5,dropin,batch > $out/${tag}_ncu_list.log 2>&1 ;;
    ncu_full)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-lz77_fast_kernel|encode_kernel|huffman_kernel}" -c ${NCU_COUNT:-4} \
        -o $out/${tag}_full${NCU_TAG} -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --only ${NCU_ONLY:-none} > $out/${tag}_ncu_full${NCU_TAG}.log 2>&1 ;;
    ncu_each)
      # one `--set full` capture per kernel (first launch that matches), each from its own short bench run.  The reports are
      # 10-16 MB each and gpurun brings back at most 64 MiB: they are condensed HERE (tools/ncu_summary.py, tools/ncu_lines.py
      # work without a GPU but need ncu + nvdisasm, which the box has) and only the text travels; NCU_KEEP names the one report to keep.
      rm -f $out/${tag}_ncu_summary.md $out/${tag}_traffic.json
      for spec in ${NCU_EACH:-lz77_fast_kernel:none huffman_kernel:none encode_kernel:none find_sync_kernel:none inflate_segments_kernel:none inflate_copy_kernel:none lz77_better_kernel:better inflate_batch_kernel:batch foreign_find_blocks_kernel:foreign foreign_decode_kernel:foreign foreign_copy_kernel:foreign foreign_window_kernel:foreign}; do
        k=${spec%%:*}; only=${spec##*:}
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o $out/${tag}_ncu_$k -f \
          python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --only $only > $out/${tag}_ncu_$k.log 2>&1
        if [ -f $out/${tag}_ncu_$k.ncu-rep ]; then
          python tools/ncu_summary.py $out/${tag}_ncu_$k.ncu-rep $out/${tag}_ncu_summary.md $out/${tag}_traffic.json > /dev/null 2>&1
          python tools/ncu_lines.py $out/${tag}_ncu_$k.ncu-rep $k 40 > $out/${tag}_lines_$k.txt 2>&1
          [ "$k" = "${NCU_KEEP:-lz77_fast_kernel}" ] || rm -f $out/${tag}_ncu_$k.ncu-rep
        fi
        tail -c 2000 $out/${tag}_ncu_$k.log > $out/${tag}_ncu_$k.log.tail; mv $out/${tag}_ncu_$k.log.tail $out/${tag}_ncu_$k.log
      done
      du -sh $out > $out/${tag}_du.txt ;;
    sanitizer)
      # memcheck and racecheck over a bounded subset (small inputs: the tools slow kernels down 10-100x); the summary lines
      # (ERROR SUMMARY / RACECHECK SUMMARY) are what profiles/ keeps
      timeout ${SAN_TIMEOUT:-480} compute-sanitizer --tool memcheck --log-file $out/${tag}_memcheck.log python -m pytest tests -m gpu -x -q -p no:cacheprovider \
        -k "${SAN_K:-fixture or quirks or edge_sizes or fuzz_single or segment_index or errors}" > $out/${tag}_memcheck_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_memcheck_pytest.log
      timeout ${SAN_TIMEOUT:-480} compute-sanitizer --tool racecheck --log-file $out/${tag}_racecheck.log python -m pytest tests -m gpu -x -q -p no:cacheprovider \
        -k "${SAN_RACE_K:-fixture or quirks or segment_index_streams or test_edge_sizes}" > $out/${tag}_racecheck_pytest.log 2>&1; echo "exit $?" >> $out/${tag}_racecheck_pytest.log
      for f in $out/${tag}_memcheck.log $out/${tag}_racecheck.log; do [ -f $f ] && { head -c 1500000 $f > $f.cut; mv $f.cut $f; }; done ;;
    probe)
      timeout 1200 python tools/probe_r02.py ${PROBE_ARGS} > $out/${tag}_probe.log 2>&1 ;;
    probe_e2e)
      for slice in 67108864 134217728 268435456; do
        B200_HOST_INFLATE_SLICE=$slice timeout 300 python tools/probe_e2e.py >> $out/${tag}_probe_e2e.log 2>&1
      done
      B200_STAGE=0 timeout 300 python tools/probe_e2e.py >> $out/${tag}_probe_e2e.log 2>&1
      B200_STAGE_THREADS=4 timeout 300 python tools/probe_e2e.py >> $out/${tag}_probe_e2e.log 2>&1
      B200_STAGE_THREADS=16 timeout 300 python tools/probe_e2e.py >> $out/${tag}_probe_e2e.log 2>&1 ;;
    env_sweep)
      # ENV_SWEEP="A=1;B=2 C=3;" : one short bench run per ;-separated setting (space-separated assignments, may be empty)
      IFS=';' read -ra settings <<< "${ENV_SWEEP}"
      for setting in "${settings[@]}" ""; do
        echo "### ${setting}" >> $out/${tag}_env_sweep.log
        env ${setting} timeout 300 python bench.py --steps 5 --warmup 2 --only none --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('compress', round(d['value'],2), d['roofline']['kernels_ms_per_step'])
print('inflate', round(d['decompress']['value'],2), d['decompress']['roofline']['kernels_ms_per_step'])" >> $out/${tag}_env_sweep.log 2>&1
      done ;;
    probe_small)
      timeout 300 python tools/probe_small.py > $out/${tag}_probe_small.log 2>&1 ;;
    probe_pcie)
      # what the PCIe link moves raw, then the pinned host-buffer calls at several inflate slice sizes
      timeout 300 python tools/probe_e2e.py 1024 pinned raw >> $out/${tag}_probe_pcie.log 2>&1
      for slice in ${PCIE_SLICES:-33554432 67108864 134217728}; do
        B200_HOST_INFLATE_SLICE=$slice timeout 300 python tools/probe_e2e.py 1024 pinned >> $out/${tag}_probe_pcie.log 2>&1
      done ;;
    bisect)
      for k in 0 1 2 3 4 5 6 7; do
        for mode in par seq; do
          echo "--- long case $k $mode" >> $out/${tag}_bisect.log
          if [ $mode = seq ]; then export B200_NO_FOREIGN_PARALLEL=1; else unset B200_NO_FOREIGN_PARALLEL; fi
          B200_DEBUG=1 timeout 40 python tools/long_case.py $k >> $out/${tag}_bisect.log 2>&1; echo "exit $?" >> $out/${tag}_bisect.log
        done
      done
      unset B200_NO_FOREIGN_PARALLEL
      timeout 600 python tools/fuzz_bisect.py >> $out/${tag}_bisect.log 2>&1; echo "exit $?" >> $out/${tag}_bisect.log ;;
    *) echo "unknown step $step" ;;
  esac
done
echo done > $out/${tag}_done.txt
