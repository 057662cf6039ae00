from .mac import MAC  # noqa: F401
from .baseline import NaiveGreedy  # noqa: F401
from .greedy_eig import GreedyEig  # noqa: F401
