#!/bin/bash
# run a command line on the GPU box with everything it prints kept in gpurun_out/<tag>/log.txt:  tools/gpu_run.sh TAG 'cmd ...'
TAG=$1; shift
mkdir -p gpurun_out/$TAG
bash -c "$*" > gpurun_out/$TAG/log.txt 2>&1
echo "rc=$? (gpurun_out/$TAG/log.txt)"
