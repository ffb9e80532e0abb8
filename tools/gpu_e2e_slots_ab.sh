python tools/pcie_probe.py
for lib in hehub_b200/libhehub_b200.so tools/_variants_slots*.so; do echo $lib; HEHUB_B200_LIB=$lib python tools/e2e_sweep.py 4096 8192 16384; done
