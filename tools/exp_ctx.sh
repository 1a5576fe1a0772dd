#!/bin/bash
# how long does the CUDA context of a drop-in executable take to open, with all GPUs of the box visible and with one?
cd "$(dirname "$0")/.."
echo "GPUs listed by nvidia-smi: $(nvidia-smi -L | wc -l); CUDA_VISIBLE_DEVICES='${CUDA_VISIBLE_DEVICES}'"
FA=tests/golden/reads.fa
for i in 1 2 3; do
  for v in "" "0"; do
    if [ -z "$v" ]; then env TRINITY_GPU_TRACE=1 trinityrnaseq_b200/bin/fastaToKmerCoverageStats --reads $FA --kmers_from_reads $FA 2>&1 >/dev/null | grep "context" | sed "s/^/all visible:  /"
    else env CUDA_VISIBLE_DEVICES=$v TRINITY_GPU_TRACE=1 trinityrnaseq_b200/bin/fastaToKmerCoverageStats --reads $FA --kmers_from_reads $FA 2>&1 >/dev/null | grep "context" | sed "s/^/one visible:  /"
    fi
  done
done
