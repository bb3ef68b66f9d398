"""Option classes with the reference's surface (``options/base_options.py``, ``aug_options.py``)."""
