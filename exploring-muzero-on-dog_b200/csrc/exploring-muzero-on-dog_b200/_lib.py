

class TttState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("board", "current_player", "reward", "done", "memory")]
