#!/usr/bin/env python
"""Prints the 64-entry table 2^(j/64) used by glmmrmcml_b200/csrc/exp_table.cuh (correctly rounded doubles)."""
import mpmath
mpmath.mp.prec = 300
for j in range(64):
    print(repr(float(mpmath.power(2, mpmath.mpf(j) / 64))))
