#!/bin/bash
bash tools/scaling_run_r2.sh 1
bash tools/scaling_run_r2.sh 2
