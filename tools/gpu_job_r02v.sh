set -x
echo ring6; timeout 150 python tools/ogemm_probe.py time_parts 2>&1 | grep '"S": 7'; timeout 150 python tools/ogemm_probe.py time_apply; timeout 150 python tools/ogemm_probe.py time_syrk
echo ring7; export VT_LIB_PATH=$PWD/vittles_b200/lib/libvittles_b200_ring7.so; timeout 150 python tools/ogemm_probe.py time_parts 2>&1 | grep '"S": 7'; timeout 150 python tools/ogemm_probe.py time_apply; timeout 150 python tools/ogemm_probe.py time_syrk
