#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled
r() { env "$@" timeout 120 python scripts/bench_emission.py 2>&1 | tail -1; }
r TAG=v1 PMG_EM_KERNEL=1
r TAG=v1_again PMG_EM_KERNEL=1
r TAG=v1_nostagger PMG_EM_KERNEL=1 PMG_EM_STAGGER=0
r TAG=v1_nostore PMG_EM_KERNEL=1 PMG_EM_NOSTORE=1
r TAG=v1_brep8 PMG_EM_KERNEL=1 PMG_EM_BREP=8
r TAG=v1_align64 PMG_EM_KERNEL=1 PMG_Y16_ALIGN=64
r TAG=v1_align64_brep8 PMG_EM_KERNEL=1 PMG_Y16_ALIGN=64 PMG_EM_BREP=8
r TAG=v1_st3 PMG_EM_KERNEL=1 PMG_EM_STAGES=3
r TAG=v2 PMG_EM_KERNEL=2
r TAG=v2_nostore PMG_EM_KERNEL=2 PMG_EM_NOSTORE=1
r TAG=v2_brep8 PMG_EM_KERNEL=2 PMG_EM_BREP=8
r TAG=v2_align64 PMG_EM_KERNEL=2 PMG_Y16_ALIGN=64
r TAG=v2_align64_brep8_nostore PMG_EM_KERNEL=2 PMG_Y16_ALIGN=64 PMG_EM_BREP=8 PMG_EM_NOSTORE=1
