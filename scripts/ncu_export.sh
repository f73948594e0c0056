#!/bin/bash
# Here (no GPU needed): turns the .ncu-rep files a profiling pass left in gpurun_out/ into the CSVs kept under profiles/.
#   scripts/ncu_export.sh <tag>
TAG=${1:-r09}
for rep in gpurun_out/prof_${TAG}_*.ncu-rep; do
  name=$(basename $rep .ncu-rep); name=${name#prof_${TAG}_}
  ncu -i $rep --page raw --csv > profiles/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $rep --page details --csv > profiles/${TAG}_${name}_details.csv 2>/dev/null
  echo "$name: $(wc -c < profiles/${TAG}_${name}_raw.csv) bytes raw"
done
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv 2>/dev/null
