from .chunking import chunk_ranges  # noqa: F401
