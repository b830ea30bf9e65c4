from scipy.linalg import *  # noqa: F401,F403
