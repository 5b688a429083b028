"""CPU oracle for the syllable-detection hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (syllable-detector-swift_b200) never does.  See oracle/oracle.c for the citation map.
"""
from .pyoracle import Oracle, OracleError, Resampler, build, build_fast, lib_path, simulator_trace  # noqa: F401
