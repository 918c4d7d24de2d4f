"""Parameter census printed by train_gp when verbose (reference directionalvi/utils/count_params.py:3-19)."""
import math


def count_params(model, likelihood):
    total = 0
    print("All parameters to learn:")
    for module in (model, likelihood):
        for name, param in module.named_parameters():
            print("     ", name)
            print("     ", param.data.shape)
            if param.requires_grad:
                total += math.prod(param.data.shape)
    print("Total number of parameters: ", total)
    return total
