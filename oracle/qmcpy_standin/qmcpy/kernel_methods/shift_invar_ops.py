from oracle.primitives import BERNOULLIPOLYSDICT, bernoulli_poly  # noqa: F401
