"""import-only stub (the reference imports `from skimage import io, color` in utils.py:5 and never uses them on the train/test path)"""
