#!/usr/bin/env python
"""Alias of tools/launch.py kept for scripts written against the reference's distgraphlaunch.py."""
import os
import runpy

runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "launch.py"), run_name="__main__")
