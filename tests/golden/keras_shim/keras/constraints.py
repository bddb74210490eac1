class Constraint(object):
    def __call__(self, w):
        return w

    def get_config(self):
        return {}
