#!/usr/bin/env python
"""Summarise registers / spills / shared memory per kernel from the build's ptxas -v logs."""
import glob, re, subprocess, sys, os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'sph-particle-simulator_b200', 'csrc', 'build')
for f in sorted(glob.glob(os.path.join(root, '*.ptxas.log'))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", txt):
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'sphb::\(anonymous namespace\)::', '', name)
        name = re.sub(r'\(.*', '', name)[:48]
        print(f"{name:48s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill={m.group(3)}/{m.group(4)} {m.group(6).strip()[:70]}")
