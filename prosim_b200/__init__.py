"""prosim_b200: B200-native closed-loop rollout path of ProSim."""
from .config import get_config  # noqa: F401
from .registry import registry  # noqa: F401


def __getattr__(name):
    # the model pulls in the native library; import it lazily so CPU-only tooling can use config / synthetic / weights
    if name == 'ProSimB200':
        from .model import ProSimB200
        return ProSimB200
    raise AttributeError(name)
