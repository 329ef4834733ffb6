import numpy


def rainbow(x):
    return numpy.zeros((len(numpy.atleast_1d(x)), 4))
