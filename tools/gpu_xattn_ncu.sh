#!/bin/bash
AB_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'xattention2' -c 2 -f -o gpurun_out/r02_xattn2 python tools/ab.py xattn > gpurun_out/r02_ncu_xattn2.log 2>&1
tail -3 gpurun_out/r02_ncu_xattn2.log
