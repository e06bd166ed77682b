"""Minimal stand-in for `gymnasium`, TEST INFRASTRUCTURE ONLY.

It exists so that `oracle/gen_golden.py` can import the read-only reference
(`/root/reference/pearl`) in a container that has neither gym nor gymnasium.
Only the attributes SURVEY.md Appendix C lists are provided.  Nothing in the
product (`pearl_b200/`) imports this.
"""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401


class Env:  # placeholder: only referenced in annotations / isinstance checks
    pass


class Wrapper(Env):
    def __init__(self, env=None):
        self.env = env


class ObservationWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


class RewardWrapper(Wrapper):
    pass


def make(*args, **kwargs):
    raise RuntimeError("gymnasium stub: no environments are available")
