class InvalidImageError(Exception):
    """Raised when an array is not a valid H x W x 3 image in [0, 255]
    (reference: ``pyvisim/_errors.py:5-10``)."""
