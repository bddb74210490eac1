def l2(l=0.01):
    return ("l2", l)
