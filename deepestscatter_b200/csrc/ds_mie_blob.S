/* Embeds deepestscatter_b200/data/mie_tables.f32 (2 x 4096 little-endian float32: mie, choppedMie;
 * data of DataGen src/Mie.cpp:8-8203, extracted by tools/extract_mie_tables.py) into the library. */
    .section .rodata
    .balign 16
    .global ds_mie_blob
    .type ds_mie_blob, @object
ds_mie_blob:
    .incbin "mie_tables.f32"
    .global ds_mie_blob_end
ds_mie_blob_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits
