"""flair-1_b200: B200-native (sm_100a) implementation of FLAIR-1's zone_detect / patch-predict hot path.

Layout mirrors the reference packages for that path only:
  zone_detect/  main, utils, slicing_job, dataset, model, compare, tiles, metrics   (src/zone_detect/*)
  flair/        main, utils, model, data_loader, tasks_utils, task_module, writer, metrics (src/flair/*)
  csrc/         hand-written CUDA kernels + the C ABI (include/flair_b200.h)
  _native.py    ctypes binding; build.py compiles libflairb200.so in-tree.
The directory name has a hyphen, so import it through the root-level alias module `flair1_b200`.
"""
__version__ = "0.1.0"
