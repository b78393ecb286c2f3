"""Run-time switches of the device backend.

noise   'host'   : action / target noise is drawn on the host from the same
                   generators the reference uses (torch's global CPU generator,
                   numpy RandomState(seed)) and copied to the device -- the
                   parity mode, bit-compatible streams ("RNG parity").
        'device' : Philox4x32-10 counter-based noise generated inside the kernels
                   (no host->device traffic); same distributions, different stream.
indices 'host'   : minibatch permutations come from the native numpy-compatible
                   MT19937 stream (bit-identical indices; with several ranks the
                   permutation is global and every rank keeps the rows it owns).
        'device' : permutations are generated on the GPU (Feistel bijection); with
                   several ranks each rank permutes ITS OWN rows and contributes
                   batch_size / world rows to every minibatch (static shapes, no host
                   work that grows with the number of GPUs).
wgrad_splits     : number of row splits of the weight-gradient kernel.
"""

noise = 'host'
indices = 'host'
wgrad_splits = 37      # 4 heavy tiles x 37 splits = 148 CTAs = one per B200 SM
