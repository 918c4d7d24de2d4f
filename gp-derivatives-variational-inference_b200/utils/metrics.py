"""Error metrics with the reference's names and argument order (directionalvi/utils/metrics.py:1-27):
Y = targets, Z = predictions, both torch tensors."""


def MSE(Y, Z):
    return ((Y - Z) ** 2).mean()


def MAE(Y, Z):
    return (Y - Z).abs().mean()


def RMSE(Y, Z):
    return MSE(Y, Z).sqrt()


def SMAE(Y, Z):
    return MAE(Y, Z) / Y.abs().mean()
