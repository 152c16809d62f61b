# usage: bash tools/prof4.sh SPEC [SPEC ...]   -- one ncu --set full capture (second skeleton launch) per perf_probe spec;
# reports are gzipped (gpurun brings back at most 64 MiB)
set -e
for spec in "$@"; do
  name=$(echo $spec | tr ':' '_')
  python tools/perf_probe.py $spec > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'skeleton|logreg' -s 1 -c 1 -o gpurun_out/prof_$name -f python tools/perf_probe.py $spec > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/plain_$name.log
done
gzip -f gpurun_out/prof_*.ncu-rep
ls -la gpurun_out/ | grep prof_
