#!/bin/bash
O=gpurun_out
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
python scripts/probe_cluster.py > $O/r2j_cluster_default.log 2>&1
SGM_LIB_PATH=$PWD/$PKG/libsgmpf_clrelaxed.so python scripts/probe_cluster.py > $O/r2j_cluster_relaxed.log 2>&1
paste $O/r2j_cluster_default.log $O/r2j_cluster_relaxed.log | cut -c1-230
python scripts/profile_target2.py paris > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/r2j_paris_launches.csv python scripts/profile_target2.py paris > $O/r2j_ncu_paris.log 2>&1
python scripts/ncu_launch_list.py $O/r2j_paris_launches.csv $O/r2j_paris_launch_list.json | tail -12
