#!/usr/bin/env python3
"""Same entry point name as the reference script; the engine is caviar_b200 (B200 kernels instead of cpptraj)."""
import sys

from caviar_b200.extractor import main

if __name__ == "__main__":
    sys.exit(main())
