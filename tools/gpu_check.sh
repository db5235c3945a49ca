#!/bin/bash
# One GPU-box visit: the GPU test suite, then bench lines of the workloads given as arguments (default c3).
# Everything lands under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
tag=${TAG:-run}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
tail -n 25 gpurun_out/${tag}_tests.log
for w in ${@:-c3}; do
	timeout 600 python bench.py --steps 20 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
	echo "bench $w rc=$?"
	python tools/show_bench.py gpurun_out/${tag}_bench_$w.json 2>/dev/null || tail -c 600 gpurun_out/${tag}_bench_$w.err
done
