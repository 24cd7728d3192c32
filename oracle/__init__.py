"""CPU oracle for the Pinball / skill-chaining hot path.  TEST INFRASTRUCTURE ONLY.

This package is the NumPy restatement of the algorithm named by the reference's
README (/root/reference/README.md:1-2 - the only file the reference ships):
Konidaris & Barto, "Skill Discovery in Continuous Reinforcement Learning Domains
using Skill Chaining".  The reference contains no code, tests or golden vectors,
so this oracle is the normative definition the CUDA path is checked against
(BASELINE.json north_star).

PARITY UNPINNED BY THE REFERENCE: there is nothing in /root/reference to pin it to.
It is pinned instead by the analytic known-answer tests in tests/test_oracle_*.py
 and by the published Philox4x32-10 known-answer vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package never does.
"""
from .philox import philox4x32, uniform01
from .pinball import PinballMap, PinballEnv, step_scalar, unpack_flags
from .fourier import FourierBasis
from .option import Option, OptionSet, logistic_features
from .agent import SkillChainAgent, AgentConfig

__all__ = [
    "philox4x32", "uniform01", "PinballMap", "PinballEnv", "step_scalar", "unpack_flags",
    "FourierBasis", "Option", "OptionSet", "logistic_features", "SkillChainAgent", "AgentConfig",
]
