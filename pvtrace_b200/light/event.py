"""Ray events; integer codes are shared with the device code (pvtrace/light/event.py:4-16)."""
from enum import Enum


class Event(Enum):
    GENERATE = 0
    REFLECT = 1
    TRANSMIT = 2
    ABSORB = 3
    NONRADIATIVE = 4
    SCATTER = 5
    EMIT = 6
    EXIT = 7
    REACT = 8
    KILL = 9
