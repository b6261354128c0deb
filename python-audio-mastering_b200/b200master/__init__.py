"""b200master -- host side of the B200-native mastering hot path.

``lib``      ctypes binding of libb200master.so (the C-ABI in include/b200_master.h)
``plan``     the reference's ``settings`` dict -> ``b200m_plan`` (filter design, numpy/scipy)
``engine``   ``Engine``: batch mastering on host or device buffers
``segment``  ``PcmSegment``: the few ``pydub.AudioSegment`` members the chain touches + WAV I/O
``synth``    deterministic synthetic programme material
"""
from .engine import Engine, get_engine, ms_framing  # noqa: F401
from .plan import make_plan, normalize_settings  # noqa: F401
