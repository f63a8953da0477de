from .clocks import ClockSampler  # noqa: F401
