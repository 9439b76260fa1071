from . import initialization  # noqa: F401

__all__ = ["initialization"]
