"""``dgl.function`` builtins used by the reference (only ``mean``; models/gnn.py:65)."""


class _Reducer:
    def __init__(self, name, msg_field, out_field):
        self.name = name
        self.msg_field = msg_field
        self.out_field = out_field


def mean(msg, out):
    return _Reducer("mean", msg, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Reducer("sum", msg, out)
