set -u
TAG=$1
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 300 gpurun_out/${TAG}_bench_default.err
bash profiles/run_profiles.sh $TAG all > gpurun_out/${TAG}_profiles.log 2>&1; tail -12 gpurun_out/${TAG}_profiles.log
