class BBCode:  # pragma: no cover - never used on the decode path
    def __init__(self, *a, **k):
        raise NotImplementedError("qldpc is not available in this image")
