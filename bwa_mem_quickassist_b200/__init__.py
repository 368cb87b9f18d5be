"""B200-native seed extension for BWA-MEM 0.7.8 (`ksw_extend` hot path).

The product is the C-ABI shared library ``libksw_b200.so`` (sources in ``csrc/``, interface in
``include/ksw_b200.h``).  This package is only the thin Python binding used by the tests, the
benchmark and ``__graft_entry__``; it contains no alignment code and has no CPU fallback: if the
CUDA library is missing or no GPU is usable, calls raise.
"""
from .ksw import (KswB200, KswB200Error, Cfg, JOB_DT, RES_DT, RJOB_DT, GJOB_DT, GRES_DT, AJOB_DT, ARES_DT, lib_path, load_library, make_cfg,
                  ksw_extend, ksw_extend2, ksw_global2, ksw_align2, extend_batch_multi, PinnedArray, pinned_copy, KswQueue)

__all__ = ["KswB200", "KswB200Error", "Cfg", "JOB_DT", "RES_DT", "RJOB_DT", "GJOB_DT", "GRES_DT", "AJOB_DT", "ARES_DT", "lib_path", "load_library", "make_cfg",
           "ksw_extend", "ksw_extend2", "ksw_global2", "ksw_align2", "extend_batch_multi", "PinnedArray", "pinned_copy", "KswQueue"]
