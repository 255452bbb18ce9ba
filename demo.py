#!/usr/bin/env python
"""Drop-in for the reference's demo.py command line (demo.py:180-270, scripts/demo_image.sh, scripts/demo_scene.sh)."""
from pixelsynth_b200.demo import main

if __name__ == "__main__":
    main()
