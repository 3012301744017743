"""``utils.common`` of the reference (utils/common.py:7-97), served by airpollution_b200.common."""
from airpollution_b200.common import AdDifProblem, Domain, Problem, backend  # noqa: F401
