#!/bin/bash
# one ncu --set full capture of the stream kernel per row dtype (1 GPU); CSV pages are produced on the box, the .ncu-rep
# is kept only when small enough for the 64 MiB return channel
mkdir -p gpurun_out
for dt in ${DTYPES:-fp32 bf16}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_stream -s 1 -c 1 -f -o /tmp/ncu_stream_${dt} \
      python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-baseline --dtype ${dt} ${EXTRA} > gpurun_out/ncu_stream_${dt}.log 2>&1
  echo "[ncu ${dt}] rc=$? $(tail -1 gpurun_out/ncu_stream_${dt}.log | cut -c1-200)"
  ncu -i /tmp/ncu_stream_${dt}.ncu-rep --page raw --csv > gpurun_out/ncu_stream_${dt}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu_stream_${dt}.ncu-rep --page source --csv > gpurun_out/ncu_stream_${dt}_source.csv 2>/dev/null
  ncu -i /tmp/ncu_stream_${dt}.ncu-rep --page details > gpurun_out/ncu_stream_${dt}_details.txt 2>/dev/null
  sz=$(stat -c %s /tmp/ncu_stream_${dt}.ncu-rep); echo "rep size $sz"
  if [ "$sz" -lt 20000000 ]; then cp /tmp/ncu_stream_${dt}.ncu-rep gpurun_out/; fi
done
