from .clip import *  # noqa: F401,F403
