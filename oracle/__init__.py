"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the Heston Monte-Carlo hot path: NumPy restatement
(heston_numpy), Philox/Box-Muller model (philox_numpy), C restatement with
explicit roundings (heston_oracle.c -> liboracle.so, loaded by c_oracle) and the
import-stub runner of the real reference (reference_runner, build container
only).  The product package cocos_b200/ never imports anything from here.
"""
