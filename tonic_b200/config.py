"""Run-time switches of the device backend.

noise   'host'   : action / target noise is drawn on the host from the same
                   generators the reference uses (torch's global CPU generator,
                   numpy RandomState(seed)) and copied to the device -- the
                   parity mode, bit-compatible streams ("RNG parity").
        'device' : Philox4x32-10 counter-based noise generated inside the kernels
                   (no host->device traffic); same distributions, different stream.
wgrad_splits     : number of row splits of the weight-gradient kernel.
"""

noise = 'host'
wgrad_splits = 37      # 4 heavy tiles x 37 splits = 148 CTAs = one per B200 SM
