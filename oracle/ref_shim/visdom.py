class Visdom:
    pass
