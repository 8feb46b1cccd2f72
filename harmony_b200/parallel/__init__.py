from .peer import PeerGroup, PeerArena  # noqa: F401
