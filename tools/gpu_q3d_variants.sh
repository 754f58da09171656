#!/bin/bash
# config 4 timing under the XL layout variants
for env in "" "OMG_B200_XL_KGLOBAL=1" "OMG_B200_CTAS=1" "OMG_B200_XL_KGLOBAL=1 OMG_B200_CTAS=1"; do
  echo "== $env"
  env $env bash tools/gpu_q3d.sh 2>&1 | grep -E "batch|config4 obst" | cut -c1-200
done
