#!/bin/bash
# Same as sweep_kernel.sh over alternative builds of the library (esp-audio-libs_b200/variants/lib_*.so).
for lib in esp-audio-libs_b200/variants/lib_*.so; do tools/sweep_kernel.sh "ESPB_LIBRARY=$PWD/$lib"; done
