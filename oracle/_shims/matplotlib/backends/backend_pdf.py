class PdfPages:
    def __init__(self, *a, **k):
        pass
