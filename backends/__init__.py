"""Dream Lab worker boundary, B200 edition.

Same module names as the reference (`backends.base`, `backends.worker_factory`,
`backends.worker_pool`) so `server/lcm_sr_server.py` keeps importing them unchanged; the new
`backends.b200_worker.B200Worker` is what the factory returns for SD1.5-class models.
"""
