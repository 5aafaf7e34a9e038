"""TEST INFRASTRUCTURE ONLY.  fmoe.gates restatement (BaseGate, NaiveGate)."""
from .base_gate import BaseGate
from .naive_gate import NaiveGate
