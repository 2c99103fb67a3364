#!/usr/bin/env python
"""``integration/run_tool.py`` with every kernel entry point replaced by its CPU emulation (tests/test_training_cpu.py:
``apply_emulation_in_this_process``) -- how the CPU test suite runs the reference's unmodified tools without a GPU."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

import test_training_cpu  # noqa: E402

test_training_cpu.apply_emulation_in_this_process()
torch.cuda.memory_allocated = lambda *a, **k: 0          # lib/core/function.py logs it
torch.nn.Module.cuda = lambda self, *a, **k: self

from integration import run_tool  # noqa: E402

run_tool.main()
