# usage: bash tools/gpu_profile.sh <tag> [kernel-regex ...]   (on the GPU box, via gpurun)
TAG=${1:-prof}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/profile_driver.py 24 2 > $OUT/driver.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_driver.py 24 1 > $OUT/ncu_launch.log 2>&1
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $OUT/$K python tools/profile_driver.py 24 1 > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT
