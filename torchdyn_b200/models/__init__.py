from .cnf import CNF, autograd_trace, hutch_trace

__all__ = ["CNF", "autograd_trace", "hutch_trace"]
