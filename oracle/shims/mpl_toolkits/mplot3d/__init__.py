class Axes3D:  # stub (amp_exit.py:7)
    pass
