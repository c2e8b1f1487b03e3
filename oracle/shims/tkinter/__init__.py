"""Import shim (test infrastructure) for the reference's stray `from tkinter.messagebox import NO`."""
