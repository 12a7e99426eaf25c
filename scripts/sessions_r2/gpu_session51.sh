# two contexts filtering different column groups on one device at the same time: does K3 of one overlap K1 of the other?
mkdir -p gpurun_out
CHEFSI_B200_GRIDSYNC=0 timeout 280 python scripts/two_context_overlap.py > gpurun_out/r2_s51_overlap_nosync.log 2>&1; tail -4 gpurun_out/r2_s51_overlap_nosync.log
CHEFSI_B200_GRIDSYNC=1 timeout 280 python scripts/two_context_overlap.py > gpurun_out/r2_s51_overlap_sync.log 2>&1; tail -4 gpurun_out/r2_s51_overlap_sync.log
