"""Process launchers: ``bpslaunch`` (one process per GPU on this host, or a
server/scheduler role) and ``dist_launcher`` (ssh fan-out from host files)."""
