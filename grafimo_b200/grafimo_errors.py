"""Exception types of the motif-scanning path (names follow src/grafimo/grafimo_errors.py of the reference
so that callers catching them keep working)."""


class NoDataFrameException(Exception):
    pass


class WrongMotifWidthException(Exception):
    pass


class WrongMotifIDException(Exception):
    pass


class WrongMotifNameException(Exception):
    pass


class NotValidMotifMatrixError(Exception):
    pass


class NotValidBGException(Exception):
    pass


class NotValidAlphabetException(Exception):
    pass


class NotValidFFException(Exception):
    pass


class FileReadError(Exception):
    pass


class FileWriteError(Exception):
    pass


class MotifFileReadError(Exception):
    pass


class MotifFileFormatError(Exception):
    pass


class BGFileError(Exception):
    pass


class MotifProcessingError(Exception):
    pass


class ValueException(Exception):
    pass


class ScoringError(Exception):
    pass


class VGError(Exception):
    pass


class SubprocessError(Exception):
    pass
