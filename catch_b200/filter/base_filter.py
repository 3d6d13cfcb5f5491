"""Filter plugin base class: the drop-in boundary.

Same contract as the reference's catch/filter/base_filter.py:37-179:
  filter(input, target_genomes=None, input_is_grouped=False, num_processes=None)
dispatches to the subclass's _filter(input[, target_genomes]).  A subclass that sets
`requires_probe_groupings = True` receives all groupings at once.  The reference forks a
multiprocessing.Pool over groupings otherwise (:111-165); a CUDA context must never cross a
fork, so here groupings are processed one after the other in-process, in the reference's
descending-size order (:128-130), which is also the order in which a single-worker reference
run consumes the `random` stream.
"""
import inspect


def set_max_num_processes_for_filter_over_groupings(max_num_processes=8):
    """Kept for interface compatibility (base_filter.py:12-30); the device path has no pool."""
    global _fg_max_num_processes
    _fg_max_num_processes = max_num_processes


set_max_num_processes_for_filter_over_groupings()


class BaseFilter:
    def filter(self, input, target_genomes=None, input_is_grouped=False, num_processes=None):
        n_params = len(inspect.signature(self._filter).parameters)
        pass_groupings = getattr(self, 'requires_probe_groupings', False) is True
        if pass_groupings:
            assert input_is_grouped is True
            return self._filter(input, target_genomes) if n_params == 2 else self._filter(input)
        if not input_is_grouped:
            return self._filter(input, target_genomes) if n_params == 2 else self._filter(input)
        order = sorted(range(len(input)), key=lambda i: len(input[i]), reverse=True)
        out = [None] * len(input)
        for i in order:
            out[i] = self._filter(input[i], target_genomes) if n_params == 2 else self._filter(input[i])
        return out

    def _filter(self, input):
        raise Exception("A subclass of BaseFilter must implement _filter(..)")
