// CPU emulation build of the image input-pipeline kernel: g++ -std=c++20 -O1 -shared -fPIC -pthread -DMPL_CPU_EMULATION
//   -I tests/dev tests/dev/preprocess_emu.cpp -o <out>.so        (tests/test_preprocess_cpu.py does this)
// Exports mpl_emu_preprocess_images(jobs, n_jobs, variant). Test infrastructure only.
#include "../../medplib_b200/csrc/preprocess.cu"
