from speech_decoding import _extend

__path__ = _extend(list(__path__), ["speech_decoding", "utils"])
