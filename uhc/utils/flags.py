"""uhc/utils/flags.py of the reference: one mutable global debug flag (scripts/train_uhc.py:30,51)."""


class Flags:
    def __init__(self):
        self.debug = False


flags = Flags()
